"""Parity-test configurations shared by oracle/make_golden.py and tests/. TEST INFRASTRUCTURE."""
import copy

from alpro_b200.configs import BASE_BERT, BASE_VIDEO  # noqa: E402,F401  (one statement of the released constants)


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
