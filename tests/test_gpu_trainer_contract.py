"""GPU: the contract between the model and an UNCHANGED reference trainer loop (run_video_retrieval.py:428-490):
optimizers that update through `p.data` (src/optimization/adamw.py:80-98), gradient accumulation
(`gradient_accumulation_steps`, run_video_retrieval.py:432-446), `zero_grad(set_to_none=False)`, hand-assigned p.grad."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import configs  # noqa: E402
from tests import helpers  # noqa: E402
from tests.test_gpu_parity import build_cuda_model, to_cuda  # noqa: E402


def _setup(seed_shift=0):
    cfg = dict(configs.GOLDEN["tiny_retrieval"])
    cfg["seed"] = cfg["seed"] + seed_shift
    spec, sd, batch = helpers.make_inputs(cfg)
    return cfg, sd, to_cuda(batch)


def _loss(out):
    return out["itc_loss"] + out["itm_loss"]


def test_data_update_is_seen_by_next_forward():
    """`p.data.add_()` does not bump `_version`; the 16-bit GEMM operand copies must follow anyway."""
    cfg, sd, batch = _setup()
    model = build_cuda_model(cfg, sd)
    out0 = model(batch)
    _loss(out0).backward()
    g = torch.Generator(device="cuda").manual_seed(1)
    with torch.no_grad():
        for p in model.parameters():
            v0 = p._version
            p.data.add_(0.02 * torch.randn(p.shape, device="cuda", generator=g))
            assert p._version == v0          # the premise of the bug
    out1 = model(batch)
    fresh = build_cuda_model(cfg, {k: v.detach().cpu() for k, v in model.state_dict().items()})
    ref = fresh(batch)
    assert abs(float(out1["itm_loss"]) - float(out0["itm_loss"])) > 1e-6
    for k in ("itc_loss", "itm_loss", "itm_scores"):
        assert torch.allclose(out1[k].detach(), ref[k].detach(), rtol=1e-6, atol=1e-7), k
    # eval forward right after a training step + .data update must not reuse the step's copies either
    _loss(out1).backward()
    with torch.no_grad():
        for p in model.parameters():
            p.data.mul_(1.01)
        model.eval()
        out2 = model(batch)
        fresh2 = build_cuda_model(cfg, {k: v.detach().cpu() for k, v in model.state_dict().items()})
        ref2 = fresh2(batch)
    assert torch.allclose(out2["itm_scores"], ref2["itm_scores"], rtol=1e-6, atol=1e-7)


def test_reference_style_optimizer_trains():
    """Three steps of a `.data`-updating optimizer: the loss of a fixed batch goes down (stale operands would freeze
    the GEMM weights while biases / LayerNorm move)."""
    cfg, sd, batch = _setup()
    model = build_cuda_model(cfg, sd)
    losses = []
    for _ in range(4):
        out = model(batch)
        loss = _loss(out)
        losses.append(float(loss))
        loss.backward()
        with torch.no_grad():
            for p in model.parameters():
                if p.grad is not None:
                    p.data.add_(p.grad, alpha=-2e-3)
                p.grad = None
    assert losses[-1] < losses[0] - 1e-3, losses


@pytest.mark.parametrize("mode", ["accumulate", "zero_in_place"])
def test_gradient_accumulation_keeps_flat_store_and_p_grad_aliased(mode):
    """with an existing p.grad the flat store that all-reduce / FusedAdamW use must hold the
    ACCUMULATED gradient and still alias p.grad."""
    cfg, sd, b0 = _setup()
    _, _, b1 = _setup(seed_shift=5)
    model = build_cuda_model(cfg, sd)

    def grads_of(batch):
        for p in model.parameters():
            p.grad = None
        _loss(model(batch)).backward()
        return {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    g0, g1 = grads_of(b0), grads_of(b1)
    for p in model.parameters():
        p.grad = None
    _loss(model(b0)).backward()
    if mode == "zero_in_place":
        for p in model.parameters():
            if p.grad is not None:
                p.grad.zero_()
        want = g1
    else:
        want = {n: g0[n] + g1[n] for n in g0}
    _loss(model(b1)).backward()
    G = model.engine.last_grads
    assert model._grads_aliased
    scale = max(float(v.abs().max()) for v in want.values())
    for n, p in model.named_parameters():
        if n not in want:
            continue
        assert p.grad.data_ptr() == G[n].data_ptr(), n                       # still one buffer
        assert float((p.grad - want[n]).abs().max()) <= 2e-3 * scale + 1e-7, n   # split-K atomics: not bit-exact


def test_hand_assigned_grad_is_flagged_not_lost():
    cfg, sd, b0 = _setup()
    model = build_cuda_model(cfg, sd)
    _loss(model(b0)).backward()
    ref = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    for p in model.parameters():
        if p.grad is not None:
            p.grad = torch.ones_like(p)         # a buffer that does not alias the flat store
    _loss(model(b0)).backward()
    assert model._grads_aliased is False
    scale = max(float(v.abs().max()) for v in ref.values())
    for n, p in model.named_parameters():
        if n in ref:
            assert float((p.grad - 1.0 - ref[n]).abs().max()) <= 2e-3 * scale + 1e-6, n
