"""GPU: train-mode regularisers (BERT hidden dropout, TimeSformer DropPath). RNG streams cannot be bit-matched with
torch's, so parity is checked by INJECTION: the masks the CUDA path drew (stateless counter hash / bernoulli factors,
kept in the step context) are fed to the CPU oracle, and losses + gradients must agree as in eval mode."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import alpro_oracle, configs  # noqa: E402
from tests import helpers  # noqa: E402
from tests.test_gpu_parity import build_cuda_model, to_cuda  # noqa: E402

DEV = "cuda"


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def test_dropout_mask_statistics_and_determinism():
    from alpro_b200 import ops
    m = torch.empty(1 << 20, device=DEV, dtype=torch.float16)
    ops.dropout_mask(m, 0.1, 1234)
    vals = torch.unique(m.float())
    assert vals.numel() == 2 and float(vals[0]) == 0.0 and abs(float(vals[1]) - 1 / 0.9) < 1e-3
    keep = float((m > 0).float().mean())
    assert abs(keep - 0.9) < 2e-3
    m2 = torch.empty_like(m)
    ops.dropout_mask(m2, 0.1, 1234)
    assert torch.equal(m, m2)                       # stateless: same (seed, index) -> same mask
    ops.dropout_mask(m2, 0.1, 1235)
    assert float((m2 != m).float().mean()) > 0.1   # different seed -> different mask
    odd = torch.empty(1001, device=DEV, dtype=torch.float16)
    ops.dropout_mask(odd, 0.5, 7)
    assert abs(float((odd > 0).float().mean()) - 0.5) < 0.08


def test_epilogue_row_scales_and_mask_multiply():
    from alpro_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(3)
    M, N, K = 300, 768, 192
    a = (torch.randn(M, K, device=DEV, generator=g) * 0.5).half()
    w = (torch.randn(N, K, device=DEV, generator=g) * 0.5).half()
    bias = torch.randn(N, device=DEV, generator=g)
    resid = torch.randn(M, N, device=DEV, generator=g)
    rsa = torch.rand(M, device=DEV, generator=g) + 0.5
    rsb = torch.rand(M, device=DEV, generator=g) + 0.5
    out = torch.empty(M, N, device=DEV)
    ops.gemm16(a, w, bias=bias, resid=resid, out32=out, row_scale=rsa, row_scale_bias=rsb)
    ref = (a.float() @ w.float().t()) * rsa[:, None] + bias[None] * rsb[:, None] + resid
    assert rel(out, ref) < 1e-4
    o16 = torch.empty(M, N, device=DEV, dtype=torch.float16)
    ops.gemm16(a, w, bias=bias, out16=o16, row_scale=rsa)
    assert rel(o16.float(), (a.float() @ w.float().t() + bias[None]) * rsa[:, None]) < 2e-3
    mask = torch.empty(M, N, device=DEV, dtype=torch.float16)
    ops.dropout_mask(mask, 0.1, 99)
    ops.gemm16(a, w, bias=bias, resid=resid, out32=out, act=ops.ACT_GELU_GRAD, aux=mask)   # (acc+bias)*mask + resid
    assert rel(out, (a.float() @ w.float().t() + bias[None]) * mask.float() + resid) < 1e-4


@pytest.mark.parametrize("M", [257, 5003])
def test_layernorm_train_hooks(M):
    from alpro_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(4)
    d = 768
    x = torch.randn(M, d, device=DEV, generator=g)
    gamma, beta = 1 + 0.1 * torch.randn(d, device=DEV, generator=g), 0.1 * torch.randn(d, device=DEV, generator=g)
    mask = torch.empty(M, d, device=DEV, dtype=torch.float16)
    ops.dropout_mask(mask, 0.1, 5)
    o32 = torch.empty(M, d, device=DEV)
    st = torch.empty(2, M, device=DEV)
    ops.layernorm_fwd(x, gamma, beta, 1e-12, out32=o32, mean=st[0], rstd=st[1], mul16=mask)
    xr = x.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (d,), gamma, beta, 1e-12)
    assert rel(o32, ref.detach() * mask.float()) < 1e-5
    dy = torch.randn(M, d, device=DEV, generator=g)
    dx = torch.empty(M, d, device=DEV)
    ops.layernorm_bwd(dy, x, st[0], st[1], gamma, dx, 0, dy_mul16=mask)
    (ref * mask.float()).backward(dy)
    assert rel(dx, xr.grad) < 2e-5
    # branch-gradient hooks: dx16 = dx * mask2 * rs ; colsum = sum_rows dx * mask2 * rsc
    mask2 = torch.empty(M, d, device=DEV, dtype=torch.float16)
    ops.dropout_mask(mask2, 0.1, 6)
    rs, rsc = torch.rand(M, device=DEV, generator=g) + 0.5, torch.rand(M, device=DEV, generator=g) + 0.5
    dx2 = torch.empty(M, d, device=DEV)
    dx16 = torch.empty(M, d, device=DEV, dtype=torch.float16)
    cs = torch.zeros(d, device=DEV)
    ops.layernorm_bwd(dy, x, st[0], st[1], gamma, dx2, 0, dx16=dx16, colsum=cs, dx16_mul16=mask2, dx16_row_scale=rs,
                      colsum_row_scale=rsc)
    xr.grad = None
    ref2 = F.layer_norm(xr, (d,), gamma, beta, 1e-12)
    ref2.backward(dy)
    assert rel(dx2, xr.grad) < 2e-5
    assert rel(dx16.float(), xr.grad * mask2.float() * rs[:, None]) < 1e-3
    assert rel(cs, (xr.grad * mask2.float() * rsc[:, None]).sum(0)) < 1e-4


def _collect_train_masks(model, cfg):
    return helpers.collect_train_masks(model, cfg)


def test_train_mode_step_matches_oracle_with_injected_masks():
    cfg = dict(configs.GOLDEN["tiny_retrieval"])
    spec, sd, batch = helpers.make_inputs(cfg)
    model = build_cuda_model(cfg, sd)
    model.train()
    torch.manual_seed(123)
    out = model(to_cuda(batch))
    tr = _collect_train_masks(model, cfg)      # before backward: the step context is released block by block
    (out["itc_loss"] + out["itm_loss"]).backward()
    assert any(d is not None for d in tr["drop_path"])          # stochastic depth really was active
    assert float((tr["emb"] == 0).float().mean()) > 0.05         # and so was dropout
    am = next(iter(tr["pos"].values()))[2]
    assert am is not None and abs(float((am == 0).float().mean()) - 0.1) < 0.02   # attention-prob dropout too
    # eval-mode result must differ (the regularisers do something) ...
    model.eval()
    out_eval = model(to_cuda(batch))
    assert abs(float(out_eval["itm_loss"]) - float(out["itm_loss"])) > 1e-5
    # ... and the oracle with the same masks must agree
    sd_o = {k: v.clone() for k, v in sd.items()}
    from alpro_b200 import synth
    for k in list(sd_o):
        c = synth.canonical_name(k)
        if c != k:
            sd_o[k] = sd_o[c]
    for v in sd_o.values():
        if v.is_floating_point():
            v.requires_grad_(True)
    ref = alpro_oracle.retrieval_forward(sd_o, cfg["bert"], cfg["vis"], batch, train=tr)
    assert out["_neg_video"].tolist() == ref["_neg_video"] and out["_neg_text"].tolist() == ref["_neg_text"]
    for k in ("itc_loss", "itm_loss", "itm_scores"):
        assert helpers.rel_err(out[k].detach().float().cpu(), ref[k].detach()) < 1e-3, k
    (ref["itc_loss"] + ref["itm_loss"]).backward()
    gmax = max(float(v.grad.abs().max()) for v in sd_o.values() if v.grad is not None)
    bad = []
    for n, p in model.named_parameters():
        r = sd_o[n].grad
        if r is None or float(r.abs().max()) < 1e-6 * gmax:
            continue
        e = helpers.rel_err(p.grad.cpu(), r)
        if e > 3e-2:
            bad.append((n, e))
    assert not bad, bad[:8]
