"""Parity-test configurations shared by oracle/make_golden.py and tests/. TEST INFRASTRUCTURE."""
import copy

BASE_BERT = {  # /root/reference/config_release/base_model.json
    "attention_probs_dropout_prob": 0.1, "hidden_act": "gelu", "hidden_dropout_prob": 0.1, "hidden_size": 768,
    "initializer_range": 0.02, "intermediate_size": 3072, "layer_norm_eps": 1e-12, "max_position_embeddings": 512,
    "model_type": "bert", "num_attention_heads": 12, "num_hidden_layers": 12, "pad_token_id": 0,
    "type_vocab_size": 2, "vocab_size": 30522, "fusion_layer": 6, "encoder_width": 768, "itc_token_type": "cls",
}
BASE_VIDEO = {  # /root/reference/config_release/timesformer_divst_8x32_224_k600.json
    "cls": "TimeSformer", "patch_size": 16, "attn_drop_rate": 0, "drop_rate": 0, "drop_path_rate": 0.1,
    "maxpool_kernel_size": 2, "use_maxpooling": False, "gradient_checkpointing": False,
}


def tiny(kind, B=2, T=2, img=64, L=8, d=192, depth=2, heads=3, bert_layers=4, fusion_layer=2, vocab=1000,
         num_entities=48, seed=0):
    """BASELINE.json configs[0]-style plumbing config: TimeSformer-tiny d=192, 2 frames, 8-token captions."""
    bert = copy.deepcopy(BASE_BERT)
    bert.update(hidden_size=d, intermediate_size=4 * d, num_attention_heads=heads, num_hidden_layers=bert_layers,
                fusion_layer=fusion_layer, vocab_size=vocab, encoder_width=d, max_position_embeddings=64)
    video = copy.deepcopy(BASE_VIDEO)
    video.update(num_frm=T, img_size=img)
    vis = dict(d=d, depth=depth, heads=heads, T=T, img=img, patch=16)
    return dict(kind=kind, B=B, T=T, img=img, L=L, bert=bert, video=video, vis=vis, num_entities=num_entities,
                seed=seed)


GOLDEN = {
    # name: config. Kept small so fixtures are a few hundred KB each.
    "tiny_retrieval": tiny("retrieval", B=3, T=2, img=64, L=8, seed=11),
    "tiny_pretrain": tiny("pretrain", B=3, T=2, img=64, L=8, seed=12),
    "tiny224_retrieval": tiny("retrieval", B=1, T=2, img=224, L=8, depth=1, bert_layers=2, fusion_layer=1, seed=13),
    "tiny_t4_retrieval": tiny("retrieval", B=2, T=4, img=48, L=12, seed=14),
}
