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


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def collect_train_masks(model, cfg):
    """Regulariser masks of the CUDA model's last train-mode forward (kept in the step context), reshaped for the
    oracle's `train=` injection (oracle/alpro_oracle.py retrieval_forward / pretrain_forward). Call BEFORE backward:
    the step context is released block by block."""
    from alpro_b200 import ops
    ctx = model.engine.last_ctx
    B = cfg["B"]
    tr = {"drop_path": [None if blk["dp"] is None else {k: v.cpu() for k, v in blk["dp"]["raw"].items()}
                        for blk in ctx["vctx"]["blocks"]]}
    h = cfg["bert"]["hidden_size"]
    L, R = ctx["L"], ctx["R"]
    nt, S_all = ctx["nt"], ctx["S_all"]
    heads = cfg["bert"]["num_attention_heads"]
    emb = ctx["ectx"]["mask"].float().cpu().view(-1, L, h)
    tr["emb"] = emb[:B]
    if nt > B:
        tr["emb_mlm"] = emb[B:2 * B]

    def amask(c, S, nseq):
        if not c["pattn"]:
            return None
        m = torch.empty(nseq, heads, S, S, device="cuda")
        ops.attn_dropout_mask(m, S, nseq, heads, c["pattn"], c["aseed"])
        return m.cpu()

    tr["text"], tr["text_mlm"] = {}, {}
    for c in ctx["tctx"]["layers"]:
        am = amask(c, L, nt)
        mo, mf = c["mo"].float().cpu().view(-1, L, h), c["mf"].float().cpu().view(-1, L, h)
        tr["text"][c["i"]] = (mo[:B], mf[:B], None if am is None else am[:B])
        if nt > B:
            tr["text_mlm"][c["i"]] = (mo[B:2 * B], mf[B:2 * B], None if am is None else am[B:2 * B])
    tr["pos"], tr["neg"], tr["mlm"] = {}, {}, {}
    for c in ctx["fctx"]["layers"]:
        a, b = c["mo"].float().cpu().view(-1, R, h), c["mf"].float().cpu().view(-1, R, h)
        am = amask(c, R, S_all)
        tr["pos"][c["i"]] = (a[:B], b[:B], None if am is None else am[:B])
        tr["neg"][c["i"]] = (a[B:3 * B], b[B:3 * B], None if am is None else am[B:3 * B])
        if S_all > 3 * B:
            tr["mlm"][c["i"]] = (a[3 * B:4 * B], b[3 * B:4 * B], None if am is None else am[3 * B:4 * B])
    if nt == B:
        for k in ("emb_mlm", "text_mlm", "mlm"):
            tr.pop(k, None)
    return tr
