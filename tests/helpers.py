"""Shared helpers for parity tests."""
import os

import numpy as np
import torch

from alpro_b200 import synth
from oracle import alpro_oracle, configs

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))


def make_inputs(cfg):
    spec = synth.model_spec(cfg["kind"], cfg["bert"], cfg["vis"], cfg["num_entities"])
    sd = synth.synth_state_dict(spec, cfg["seed"])
    batch = synth.synth_batch(cfg["kind"], cfg["B"], cfg["T"], cfg["img"], cfg["L"], cfg["bert"]["vocab_size"],
                              seed=cfg["seed"], num_entities=cfg["num_entities"])
    return spec, sd, batch


def oracle_run(cfg, sd, batch, requires_grad=False):
    sd = {k: v.clone() for k, v in sd.items()}
    # tied parameters share one leaf (synth gives identical values; make them the same tensor for autograd)
    for k in list(sd):
        c = synth.canonical_name(k)
        if c != k:
            sd[k] = sd[c]
    if requires_grad:
        for k, v in sd.items():
            if v.is_floating_point() and "prompter." not in k and not k.endswith("prompt_feat"):
                v.requires_grad_(True)
    fwd = alpro_oracle.retrieval_forward if cfg["kind"] == "retrieval" else alpro_oracle.pretrain_forward
    out = fwd(sd, cfg["bert"], cfg["vis"], batch)
    return sd, out


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
