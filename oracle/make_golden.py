"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on seeded synthetic inputs.
TEST INFRASTRUCTURE. Run in the build container only:   python -m oracle.make_golden

Each fixture stores: the reference outputs (losses, logits, embeddings), the hard-negative indices drawn (argmax rule,
see ref_harness.FixedNegatives), gradient L2 norms of every parameter and a few full gradient tensors. Weights and
inputs are NOT stored: they are regenerated bit-exactly from (name, shape, seed) by alpro_b200.synth.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from alpro_b200 import synth  # noqa: E402
from oracle import configs, ref_harness  # noqa: E402

FULL_GRADS = ["temp", "vision_proj.weight", "itm_head.weight", "visual_encoder.model.time_embed",
              "visual_encoder.model.cls_token", "visual_encoder.model.blocks.0.temporal_fc.bias",
              "visual_encoder.model.blocks.0.norm1.weight", "text_encoder.bert.embeddings.LayerNorm.weight",
              "text_encoder.bert.encoder.layer.0.attention.self.query.bias", "mpm_head.2.bias"]


def build(cfg):
    model = ref_harness.build_reference_model(cfg["kind"], cfg["bert"], cfg["video"],
                                              vis_dims=(cfg["vis"]["d"], cfg["vis"]["depth"], cfg["vis"]["heads"]),
                                              num_entities=cfg["num_entities"])
    spec = synth.model_spec(cfg["kind"], cfg["bert"], cfg["vis"], cfg["num_entities"])
    ref_keys = list(model.state_dict().keys())
    assert ref_keys == list(spec.keys()) or set(ref_keys) == set(spec.keys()), \
        (set(ref_keys) ^ set(spec.keys()))
    for k, v in model.state_dict().items():
        assert tuple(v.shape) == tuple(spec[k]), (k, v.shape, spec[k])
    sd = synth.synth_state_dict(spec, cfg["seed"])
    model.load_state_dict(sd, strict=True)
    ref_harness.tie_mlm_decoder(model)
    model.eval()
    return model, sd


def run_reference(cfg):
    model, sd = build(cfg)
    batch = synth.synth_batch(cfg["kind"], cfg["B"], cfg["T"], cfg["img"], cfg["L"], cfg["bert"]["vocab_size"],
                              seed=cfg["seed"], num_entities=cfg["num_entities"])
    for p in model.parameters():
        p.grad = None
    with ref_harness.FixedNegatives() as fn:
        out = model(batch)
    loss = sum(v for k, v in out.items() if k.endswith("_loss") and v is not None)
    loss.backward()
    res = {}
    for k, v in out.items():
        if torch.is_tensor(v):
            res["out." + k] = v.detach().numpy()
    res["neg_drawn"] = np.asarray(fn.drawn, dtype=np.int64)
    names, norms = [], []
    for n, p in model.named_parameters():
        g = p.grad
        names.append(n)
        norms.append(0.0 if g is None else float(g.double().norm()))
        if n in FULL_GRADS and g is not None:
            res["grad." + n] = g.detach().numpy()
    res["grad_names"] = np.asarray(names)
    res["grad_norms"] = np.asarray(norms, dtype=np.float64)
    # intermediate activations through the reference's own sub-modules
    with torch.no_grad():
        ve = model.visual_encoder.forward_features(batch["visual_inputs"].transpose(1, 2), return_all_tokens=True)
        res["act.video_embeds"] = ve.numpy()
        te = model.text_encoder.bert(batch["text_input_ids"], attention_mask=batch["text_input_mask"],
                                     return_dict=True, mode="text").last_hidden_state
        res["act.text_embeds"] = te.numpy()
        tok = model.visual_encoder.model.forward_features(batch["visual_inputs"].transpose(1, 2), return_all_tokens=True)
        res["act.video_tokens_head"] = tok[:, : 1 + 2 * cfg["T"]].numpy()   # token-order check: cls + first patches
        if cfg["kind"] == "retrieval":
            inf = model.forward_inference({"visual_inputs": batch["visual_inputs"][:1],
                                           "text_input_ids": batch["text_input_ids"],
                                           "text_input_mask": batch["text_input_mask"]})
            res["inf.logits"] = inf["logits"].numpy()
            res["inf.itc_scores"] = inf["itc_scores"].numpy()
    return res


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.manual_seed(0)
    for name, cfg in configs.GOLDEN.items():
        res = run_reference(cfg)
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **res)
        print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in res.items() if k.startswith("out.")},
              os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
