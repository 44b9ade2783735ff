"""GPU parity at the BENCHED model: 12 TimeSformer blocks (d=768, 12 heads), 6 text + 6 fusion BERT layers, 8 x 224^2
frames (1569 tokens per clip), 40-token captions, vocab 30522, 1000 entities — the AlproForPretrain step of bench.py
(reference: src/modeling/alpro_models.py:79-183) on B=4 clips against the CPU oracle, in eval mode and in train mode
with the CUDA path's regulariser masks injected into the oracle, for fp16 and bf16 operands.

Tolerances (north_star: "logits/loss within 1e-3 relative fp tolerance"):
  * losses: |a-b| / |b| < 1e-3
  * logits tensors (itm_scores, mlm_scores, mpm_logits): max|a-b| / max|b| < 1e-3 for fp16 operands (the default and
    benched format); bf16 operands carry 8 mantissa bits instead of 11 and are held to 1.2e-2 on logits / 2e-3 on
    losses (measured: 3e-3 .. 8.2e-3 and <= 9.6e-4; recorded in DESIGN.md)
  * hard-negative indices, labels: bit-exact
  * parameter gradients: max|a-b| / max|b| per tensor, plus the relative L2 error
Every error is printed so the run log documents the measured deviation of the benched configuration.
"""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import alpro_oracle, configs  # noqa: E402
from tests import helpers  # noqa: E402
from tests.test_gpu_parity import build_cuda_model, to_cuda  # noqa: E402

# measured on B200 (profiles/r02_full_depth_parity.jsonl): fp16 losses <= 7e-5, logits <= 7.4e-4, gradients <= 7e-3
# (mpm_head 4.5e-2); bf16 losses <= 9.6e-4, logits <= 8.2e-3, gradients <= 5.3e-2 (mpm_head 1.9e-1)
TOL = {torch.float16: dict(loss=1e-3, logits=1e-3, emb=2e-3, labels=3e-3, grad=2e-2),
       torch.bfloat16: dict(loss=2e-3, logits=1.2e-2, emb=1.5e-2, labels=3e-2, grad=1e-1)}
LOSSES = ("itc_loss", "itm_loss", "mlm_loss", "mpm_loss")
LOGITS = ("itm_scores", "mlm_scores", "mpm_logits")
_REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "full_depth_parity.jsonl")


def full_cfg(B=4, seed=31):
    cfg = configs.tiny("pretrain", B=B, T=8, img=224, L=40, d=768, depth=12, heads=12, bert_layers=12, fusion_layer=6,
                       vocab=30522, num_entities=1000, seed=seed)
    cfg["bert"]["max_position_embeddings"] = 512
    return cfg


_cache = {}


def _inputs():
    if "in" not in _cache:
        cfg = full_cfg()
        _cache["in"] = (cfg,) + helpers.make_inputs(cfg)
    return _cache["in"]


def _oracle(cfg, sd, batch, train=None):
    sd_o = {k: v.clone() for k, v in sd.items()}
    from alpro_b200 import synth
    for k in list(sd_o):
        c = synth.canonical_name(k)
        if c != k:
            sd_o[k] = sd_o[c]
    for k, v in sd_o.items():
        if v.is_floating_point() and "prompter." not in k:
            v.requires_grad_(True)
    torch.set_num_threads(max(1, min(32, os.cpu_count() or 1)))
    out = alpro_oracle.pretrain_forward(sd_o, cfg["bert"], cfg["vis"], batch, train=train)
    sum(out[k] for k in LOSSES).backward()
    return sd_o, out


def _compare(tag, dtype, model, out, sd_o, ref):
    tol = TOL[dtype]
    assert out["_neg_video"].tolist() == ref["_neg_video"] and out["_neg_text"].tolist() == ref["_neg_text"]
    assert out["itm_labels"].cpu().tolist() == ref["itm_labels"].tolist()
    assert out["_mpm_ignore"].cpu().tolist() == ref["_mpm_ignore"].tolist()
    errs = {}
    for k in LOSSES:
        a, b = float(out[k]), float(ref[k])
        errs[k] = abs(a - b) / abs(b)
    for k in LOGITS + ("mpm_labels",):
        errs[k] = helpers.rel_err(out[k].detach().float().cpu(), ref[k].detach())
        errs[k + ".l2"] = helpers.rel_l2(out[k].detach().float().cpu(), ref[k].detach())
    errs["video_embeds"] = helpers.rel_err(out["_video_embeds"].cpu(), ref["_video_embeds"].detach())
    errs["text_embeds"] = helpers.rel_err(out["_text_embeds"].cpu(), ref["_text_embeds"].detach())
    gmax = max(float(v.grad.abs().max()) for v in sd_o.values() if v.grad is not None)
    gerrs = []
    for n, p in model.named_parameters():
        r = sd_o[n].grad if n in sd_o else None
        if n.startswith("prompter.") or r is None or float(r.abs().max()) < 1e-6 * gmax:
            continue
        assert torch.isfinite(p.grad).all(), n
        gerrs.append((helpers.rel_err(p.grad.cpu(), r), helpers.rel_l2(p.grad.cpu(), r), n))
    gerrs.sort(reverse=True)
    rec = dict(tag=tag, dtype=str(dtype), errs={k: float(f"{v:.3e}") for k, v in errs.items()},
               worst_grads=[(n, float(f"{e:.3e}"), float(f"{l2:.3e}")) for e, l2, n in gerrs[:8]],
               median_grad_err=float(f"{gerrs[len(gerrs) // 2][0]:.3e}"), n_grads=len(gerrs))
    print("FULL-DEPTH", json.dumps(rec))
    try:
        os.makedirs(os.path.dirname(_REPORT), exist_ok=True)
        with open(_REPORT, "a") as f:
            f.write(json.dumps(rec) + "\n")
    except OSError:
        pass
    for k in LOSSES:
        assert errs[k] < tol["loss"], (k, errs[k])
    for k in LOGITS:
        assert errs[k] < tol["logits"], (k, errs[k])
    assert errs["mpm_labels"] < tol["labels"], errs["mpm_labels"]
    assert errs["video_embeds"] < tol["emb"] and errs["text_embeds"] < tol["emb"], errs
    # mpm_head gradients are a difference of two near-uniform 1000-way distributions (see test_gpu_parity)
    bad = [(n, e) for e, l2, n in gerrs if e > (4 * tol["grad"] if n.startswith("mpm_head.") else tol["grad"])]
    assert not bad, bad[:8]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16], ids=["fp16", "bf16"])
def test_full_depth_eval_vs_oracle(dtype):
    cfg, spec, sd, batch = _inputs()
    if "eval" not in _cache:
        _cache["eval"] = _oracle(cfg, sd, batch)
    sd_o, ref = _cache["eval"]
    model = build_cuda_model(cfg, sd).set_compute_dtype(dtype)
    from alpro_b200.engine import argmax_sampler
    model.engine.sampler = argmax_sampler
    out = model(to_cuda(batch))
    sum(out[k] for k in LOSSES).backward()
    torch.cuda.synchronize()
    _compare("eval", dtype, model, out, sd_o, ref)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16], ids=["fp16", "bf16"])
def test_full_depth_train_mode_vs_oracle_injected_masks(dtype):
    """The benched mode: .train() (BERT hidden + attention-probability dropout 0.1, DropPath 0.1), pretrain losses."""
    cfg, spec, sd, batch = _inputs()
    cfg = dict(cfg)
    model = build_cuda_model(cfg, sd).set_compute_dtype(dtype)
    from alpro_b200.engine import argmax_sampler
    model.engine.sampler = argmax_sampler
    model.train()
    torch.manual_seed(321)
    out = model(to_cuda(batch))
    tr = helpers.collect_train_masks(model, cfg)
    assert any(d is not None for d in tr["drop_path"]) and "mlm" in tr and "emb_mlm" in tr
    sum(out[k] for k in LOSSES).backward()
    torch.cuda.synchronize()
    sd_o, ref = _oracle(cfg, sd, batch, train=tr)
    _compare("train", dtype, model, out, sd_o, ref)
