"""Deterministic synthetic weights and batches (there is no network for checkpoints or datasets).

Weights are a pure function of (parameter name, shape, seed) through numpy's PCG64 stream, so the reference model, the
oracle and the CUDA model can all be loaded with bit-identical tensors without shipping checkpoint files.
Batch shapes/keys follow the reference collators (src/datasets/dataset_pretrain_sparse.py:252-264,
src/datasets/dataset_video_retrieval.py:130-137) as summarised in SURVEY.md §8(b)/(d).
"""
import zlib
from collections import OrderedDict

import numpy as np
import torch


# ----------------------------------------------------------------------------------------------- state_dict schema
def visual_spec(prefix, d, depth, T, n_patches, patch, num_classes=400, mlp_ratio=4):
    s = OrderedDict()
    p = prefix
    s[p + "cls_token"] = (1, 1, d)
    s[p + "pos_embed"] = (1, n_patches + 1, d)
    s[p + "time_embed"] = (1, T, d)
    s[p + "patch_embed.proj.weight"] = (d, 3, patch, patch)
    s[p + "patch_embed.proj.bias"] = (d,)
    for i in range(depth):
        b = f"{p}blocks.{i}."
        s[b + "norm1.weight"] = (d,)
        s[b + "norm1.bias"] = (d,)
        s[b + "attn.qkv.weight"] = (3 * d, d)
        s[b + "attn.qkv.bias"] = (3 * d,)
        s[b + "attn.proj.weight"] = (d, d)
        s[b + "attn.proj.bias"] = (d,)
        s[b + "temporal_norm1.weight"] = (d,)
        s[b + "temporal_norm1.bias"] = (d,)
        s[b + "temporal_attn.qkv.weight"] = (3 * d, d)
        s[b + "temporal_attn.qkv.bias"] = (3 * d,)
        s[b + "temporal_attn.proj.weight"] = (d, d)
        s[b + "temporal_attn.proj.bias"] = (d,)
        s[b + "temporal_fc.weight"] = (d, d)
        s[b + "temporal_fc.bias"] = (d,)
        s[b + "norm2.weight"] = (d,)
        s[b + "norm2.bias"] = (d,)
        s[b + "mlp.fc1.weight"] = (mlp_ratio * d, d)
        s[b + "mlp.fc1.bias"] = (mlp_ratio * d,)
        s[b + "mlp.fc2.weight"] = (d, mlp_ratio * d)
        s[b + "mlp.fc2.bias"] = (d,)
    s[p + "norm.weight"] = (d,)
    s[p + "norm.bias"] = (d,)
    s[p + "head.weight"] = (num_classes, d)
    s[p + "head.bias"] = (num_classes,)
    return s


def bert_spec(prefix, cfg):
    h, ff, V = cfg["hidden_size"], cfg["intermediate_size"], cfg["vocab_size"]
    s = OrderedDict()
    e = prefix + "bert.embeddings."
    s[e + "position_ids"] = (1, cfg["max_position_embeddings"])
    s[e + "word_embeddings.weight"] = (V, h)
    s[e + "position_embeddings.weight"] = (cfg["max_position_embeddings"], h)
    s[e + "token_type_embeddings.weight"] = (cfg["type_vocab_size"], h)
    s[e + "LayerNorm.weight"] = (h,)
    s[e + "LayerNorm.bias"] = (h,)
    for i in range(cfg["num_hidden_layers"]):
        l = f"{prefix}bert.encoder.layer.{i}."
        for n in ("query", "key", "value"):
            s[l + f"attention.self.{n}.weight"] = (h, h)
            s[l + f"attention.self.{n}.bias"] = (h,)
        s[l + "attention.output.dense.weight"] = (h, h)
        s[l + "attention.output.dense.bias"] = (h,)
        s[l + "attention.output.LayerNorm.weight"] = (h,)
        s[l + "attention.output.LayerNorm.bias"] = (h,)
        s[l + "intermediate.dense.weight"] = (ff, h)
        s[l + "intermediate.dense.bias"] = (ff,)
        s[l + "output.dense.weight"] = (h, ff)
        s[l + "output.dense.bias"] = (h,)
        s[l + "output.LayerNorm.weight"] = (h,)
        s[l + "output.LayerNorm.bias"] = (h,)
    c = prefix + "cls.predictions."
    s[c + "bias"] = (V,)
    s[c + "transform.dense.weight"] = (h, h)
    s[c + "transform.dense.bias"] = (h,)
    s[c + "transform.LayerNorm.weight"] = (h,)
    s[c + "transform.LayerNorm.bias"] = (h,)
    s[c + "decoder.weight"] = (V, h)     # tied to word_embeddings.weight
    s[c + "decoder.bias"] = (V,)         # same Parameter as cls.predictions.bias (xbert.py:677)
    return s


def base_spec(prefix, bert_cfg, vis):
    """vis = dict(d, depth, heads, T, img, patch)."""
    s = OrderedDict()
    s[prefix + "temp"] = ()
    n_patches = (vis["img"] // vis["patch"]) ** 2
    s.update(visual_spec(prefix + "visual_encoder.model.", vis["d"], vis["depth"], vis["T"], n_patches, vis["patch"]))
    s.update(bert_spec(prefix + "text_encoder.", bert_cfg))
    s[prefix + "vision_proj.weight"] = (256, vis["d"])
    s[prefix + "vision_proj.bias"] = (256,)
    s[prefix + "text_proj.weight"] = (256, bert_cfg["hidden_size"])
    s[prefix + "text_proj.bias"] = (256,)
    s[prefix + "itm_head.weight"] = (2, bert_cfg["hidden_size"])
    s[prefix + "itm_head.bias"] = (2,)
    return s


def model_spec(kind, bert_cfg, vis, num_entities=None):
    """Ordered name->shape map equal to the reference model's state_dict()."""
    s = base_spec("", bert_cfg, vis)
    h = bert_cfg["hidden_size"]
    if kind == "prompter":
        s["video_prompt_feat"] = (num_entities, 256)
        s["image_prompt_feat"] = (num_entities, 256)
    elif kind == "pretrain":
        p = base_spec("prompter.", bert_cfg, vis)
        p["prompter.video_prompt_feat"] = (num_entities, 256)
        p["prompter.image_prompt_feat"] = (num_entities, 256)
        s.update(p)
        s["mpm_head.0.weight"] = (2 * h, h)
        s["mpm_head.0.bias"] = (2 * h,)
        s["mpm_head.2.weight"] = (num_entities, 2 * h)
        s["mpm_head.2.bias"] = (num_entities,)
    elif kind != "retrieval":
        raise ValueError(kind)
    return s


TIED = {  # alias -> canonical (same storage in the reference)
    "cls.predictions.decoder.weight": "bert.embeddings.word_embeddings.weight",
    "cls.predictions.decoder.bias": "cls.predictions.bias",
}


def canonical_name(name):
    for alias, canon in TIED.items():
        if name.endswith(alias):
            return name[: -len(alias)] + canon
    return name


# ----------------------------------------------------------------------------------------------- values
def _rng(name, seed):
    return np.random.default_rng([seed, zlib.crc32(name.encode())])


def synth_tensor(name, shape, seed=0):
    name = canonical_name(name)
    r = _rng(name, seed)
    leaf = name.split(".")[-1]
    if name.endswith("position_ids"):
        return torch.arange(shape[1], dtype=torch.int64).unsqueeze(0)
    if name.endswith("temp"):
        return torch.tensor(0.07, dtype=torch.float32)
    if name.endswith("prompt_feat"):
        return torch.from_numpy(r.random(shape, dtype=np.float32))  # torch.rand in the reference (alpro_models.py:396)
    is_norm = ("norm" in name.lower().split(".")[-2]) if len(name.split(".")) > 1 else False
    if is_norm and leaf == "weight":
        v = 1.0 + 0.1 * r.standard_normal(shape, dtype=np.float32)
    elif is_norm and leaf == "bias":
        v = 0.1 * r.standard_normal(shape, dtype=np.float32)
    else:
        v = 0.02 * r.standard_normal(shape, dtype=np.float32)
    return torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))


def synth_state_dict(spec, seed=0):
    return OrderedDict((k, synth_tensor(k, tuple(shp), seed)) for k, shp in spec.items())


# ----------------------------------------------------------------------------------------------- batches
def synth_batch(kind, B, T, img, L, vocab, seed=0, min_len=None, num_entities=None):
    """Synthetic batch with the reference collators' keys. Token ids: [CLS]=101 first, [SEP]=102 last real position,
    zero padding; lengths ~U{min_len..L}. For small vocabularies the special ids are folded into range."""
    r = np.random.default_rng([seed, 7919])
    CLS, SEP, MASK = (101, 102, 103) if vocab > 200 else (1, 2, 3)
    lo = 1000 if vocab > 2000 else 4
    min_len = min_len or max(3, L // 5)
    ids = np.zeros((B, L), dtype=np.int64)
    mask = np.zeros((B, L), dtype=np.int64)
    for b in range(B):
        n = int(r.integers(min_len, L + 1))
        ids[b, :n] = r.integers(lo, vocab, size=n)
        ids[b, 0] = CLS
        ids[b, n - 1] = SEP
        mask[b, :n] = 1
    batch = {
        "visual_inputs": torch.from_numpy(r.standard_normal((B, T, 3, img, img), dtype=np.float32)),
        "text_input_ids": torch.from_numpy(ids),
        "text_input_mask": torch.from_numpy(mask),
    }
    if kind == "pretrain":
        mlm_ids = ids.copy()
        labels = np.full((B, L), -100, dtype=np.int64)
        for b in range(B):
            n = int(mask[b].sum())
            cand = np.arange(1, max(2, n - 1))
            k = max(1, int(round(0.15 * len(cand))))
            pick = r.choice(cand, size=k, replace=False)
            labels[b, pick] = ids[b, pick]
            mlm_ids[b, pick] = MASK
        g = img // 16
        mpm = np.ones((B, g, g), dtype=np.float32)
        for b in range(B):
            hh = int(r.integers(max(1, g // 2), g + 1))
            ww = max(1, min(g, int(round(0.4 * g * g / hh))))
            y0 = int(r.integers(0, g - hh + 1))
            x0 = int(r.integers(0, g - ww + 1))
            mpm[b, y0:y0 + hh, x0:x0 + ww] = 0.0
        batch.update(
            mlm_text_input_ids=torch.from_numpy(mlm_ids),
            mlm_labels=torch.from_numpy(labels),
            mpm_mask=torch.from_numpy(mpm),
            crop_visual_inputs=torch.from_numpy(r.standard_normal((B, T, 3, img, img), dtype=np.float32)),
            context_visual_inputs=torch.from_numpy(r.standard_normal((B, T, 3, img, img), dtype=np.float32)),
            type="video",
        )
    return batch
