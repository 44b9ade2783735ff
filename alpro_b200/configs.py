"""Model constants of the released ALPRO configuration (there is no network / config download on the GPU box).

BASE_BERT  = /root/reference/config_release/base_model.json (+ the attributes the trainers add: fusion_layer,
             encoder_width, itc_token_type)
BASE_VIDEO = /root/reference/config_release/timesformer_divst_8x32_224_k600.json
TimeSformer-B/16 dimensions are hard-coded in the reference (src/modeling/timesformer/vit.py:445-462).
"""
BASE_BERT = {
    "attention_probs_dropout_prob": 0.1, "hidden_act": "gelu", "hidden_dropout_prob": 0.1, "hidden_size": 768,
    "initializer_range": 0.02, "intermediate_size": 3072, "layer_norm_eps": 1e-12, "max_position_embeddings": 512,
    "model_type": "bert", "num_attention_heads": 12, "num_hidden_layers": 12, "pad_token_id": 0,
    "type_vocab_size": 2, "vocab_size": 30522, "fusion_layer": 6, "encoder_width": 768, "itc_token_type": "cls",
}
BASE_VIDEO = {
    "cls": "TimeSformer", "patch_size": 16, "attn_drop_rate": 0, "drop_rate": 0, "drop_path_rate": 0.1,
    "maxpool_kernel_size": 2, "use_maxpooling": False, "gradient_checkpointing": False,
}
TIMESFORMER_B16 = {"d": 768, "depth": 12, "heads": 12, "patch": 16}
