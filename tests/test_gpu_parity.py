"""GPU parity tests proper: the CUDA model (through the reference-facing nn.Module API and the C-ABI underneath)
against (a) the golden vectors produced by the unmodified reference and (b) the CPU oracle on the same seeded inputs.

Tolerances (north_star): losses / logits within 1e-3 relative to the tensor scale; hard-negative indices, labels and
token order bit-exact. Gradients: 1e-2 relative on parameter-gradient norms (they pass through fp16 operands twice).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import configs  # noqa: E402
from tests import helpers  # noqa: E402

FWD_TOL = 1e-3
EMB_TOL = 2e-3     # full [B,197,d] embedding tensors, max-abs / max-abs
GRAD_TOL = 1e-2


def build_cuda_model(cfg, sd):
    from alpro_b200 import modeling
    from alpro_b200.engine import argmax_sampler
    v = dict(cfg["video"])
    v.update(embed_dim=cfg["vis"]["d"], depth=cfg["vis"]["depth"], num_heads=cfg["vis"]["heads"])
    b = dict(cfg["bert"])
    b["num_entities"] = cfg["num_entities"]
    cls = modeling.AlproForVideoTextRetrieval if cfg["kind"] == "retrieval" else modeling.AlproForPretrain
    m = cls(b, v)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    m.engine.sampler = argmax_sampler
    return m


def to_cuda(batch):
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}


@pytest.mark.parametrize("name", list(configs.GOLDEN))
def test_forward_backward_vs_reference_golden(name):
    cfg = configs.GOLDEN[name]
    gold = helpers.load_golden(name)
    spec, sd, batch = helpers.make_inputs(cfg)
    model = build_cuda_model(cfg, sd)
    out = model(to_cuda(batch))
    torch.cuda.synchronize()
    # index work: bit-exact
    drawn = out["_neg_video"].tolist() + out["_neg_text"].tolist()
    assert drawn == gold["neg_drawn"].tolist()
    assert out["itm_labels"].cpu().tolist() == gold["out.itm_labels"].tolist()
    errs = {}
    for k in ("itc_loss", "itm_loss", "itm_scores", "mlm_loss", "mlm_scores", "mpm_loss", "mpm_logits", "mpm_labels"):
        if "out." + k in gold:
            errs[k] = helpers.rel_err(out[k].detach().float().cpu(), gold["out." + k])
    errs["video_embeds"] = helpers.rel_err(out["_video_embeds"].cpu(), gold["act.video_embeds"])
    errs["text_embeds"] = helpers.rel_err(out["_text_embeds"].cpu(), gold["act.text_embeds"])
    print(name, {k: f"{v:.2e}" for k, v in errs.items()})
    for k, v in errs.items():
        assert v < (EMB_TOL if k.endswith("embeds") else FWD_TOL), (k, v)
    # backward through the reference-style API: sum of losses, loss.backward(), p.grad
    loss = sum(v for k, v in out.items() if k.endswith("_loss") and v is not None)
    loss.backward()
    torch.cuda.synchronize()
    gnorm = dict(zip(gold["grad_names"].tolist(), gold["grad_norms"].tolist()))
    grads = {n: p.grad for n, p in model.named_parameters()}
    worst = ("", 0.0)
    checked = 0
    for n, ref in gnorm.items():
        if n.startswith("prompter."):
            assert grads.get(n) is None or float(grads[n].abs().max()) == 0.0
            continue
        gr = grads.get(n)
        mine = 0.0 if gr is None else float(gr.double().norm())
        if ref < 1e-7:
            assert mine < 1e-5, (n, mine)
            continue
        e = abs(mine - ref) / ref
        if e > worst[1]:
            worst = (n, e)
        checked += 1
    print(name, "worst grad-norm rel err", worst, "checked", checked)
    assert worst[1] < GRAD_TOL, worst
    for k in gold:
        if k.startswith("grad.") and not k.startswith("grad.prompter"):
            n = k[5:]
            e = helpers.rel_err(grads[n].cpu(), gold[k])
            assert e < 2 * GRAD_TOL, (n, e)


def test_matches_oracle_all_param_grads():
    """Every parameter gradient (not only norms) against the CPU oracle's autograd on the same inputs."""
    cfg = configs.GOLDEN["tiny_pretrain"]
    spec, sd, batch = helpers.make_inputs(cfg)
    sd_g, oout = helpers.oracle_run(cfg, sd, batch, requires_grad=True)
    sum(v for k, v in oout.items() if k.endswith("_loss") and v is not None).backward()
    model = build_cuda_model(cfg, sd)
    out = model(to_cuda(batch))
    sum(v for k, v in out.items() if k.endswith("_loss") and v is not None).backward()
    bad = []
    gmax = max(float(v.grad.abs().max()) for v in sd_g.values() if v.grad is not None)
    for n, p in model.named_parameters():
        if n.startswith("prompter.") or n.endswith("head.weight") and "visual_encoder" in n or n.endswith("head.bias") and "visual_encoder" in n:
            continue
        ref = sd_g[n].grad
        if ref is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, n
            continue
        if float(ref.abs().max()) < 1e-6 * gmax:
            # mathematically zero gradient (e.g. attention key bias: softmax is shift-invariant); only noise on both sides
            assert float(p.grad.abs().max()) < 1e-4 * gmax, n
            continue
        e = helpers.rel_err(p.grad.cpu(), ref)
        if e > 3e-2:
            bad.append((n, e))
    assert not bad, bad[:10]


def test_inference_matches_golden():
    cfg = configs.GOLDEN["tiny_retrieval"]
    gold = helpers.load_golden("tiny_retrieval")
    spec, sd, batch = helpers.make_inputs(cfg)
    model = build_cuda_model(cfg, sd)
    b1 = to_cuda({"visual_inputs": batch["visual_inputs"][:1], "text_input_ids": batch["text_input_ids"],
                  "text_input_mask": batch["text_input_mask"]})
    out = model.forward_inference(b1)
    assert helpers.rel_err(out["logits"].cpu(), gold["inf.logits"]) < FWD_TOL
    assert helpers.rel_err(out["itc_scores"].cpu(), gold["inf.itc_scores"]) < FWD_TOL


def test_token_order_bit_exact_on_gpu():
    """pos/time-tagged tokens must come out as [cls, (n0,t0..), (n1,t0..), ...]."""
    from alpro_b200 import ops
    B, T, N, d = 2, 4, 9, 192
    proj = torch.zeros(B * (1 + N * T), d, device="cuda")
    pos = (torch.arange(N + 1, device="cuda").float() * 100).view(N + 1, 1).expand(N + 1, d).contiguous()
    tim = torch.arange(T, device="cuda").float().view(T, 1).expand(T, d).contiguous()
    x = torch.empty_like(proj)
    ops.vit_embed_fwd(proj, torch.zeros(d, device="cuda"), pos, tim, x, B, N, T, d)
    want = [0.0] + [100.0 * (n + 1) + t for n in range(N) for t in range(T)]
    assert x.view(B, -1, d)[1, :, 0].tolist() == want


def test_no_cpu_fallback():
    cfg = configs.GOLDEN["tiny_retrieval"]
    spec, sd, batch = helpers.make_inputs(cfg)
    from alpro_b200 import modeling
    v = dict(cfg["video"]); v.update(embed_dim=192, depth=2, num_heads=3)
    m = modeling.AlproForVideoTextRetrieval(dict(cfg["bert"]), v)
    with pytest.raises(RuntimeError):
        m(batch)


def test_full_width_shapes_vs_oracle():
    """Real layer widths and sequence lengths of the bench workload (d=768/12 heads, 8x224^2 -> 1569 tokens per clip,
    197-key spatial attention, L=40, 237-token fusion sequences, vocab 30522, 1000 entities) on a shallow stack
    (2 TimeSformer blocks, 2+2 BERT layers) so that the CPU oracle finishes in seconds: pretrain step, fwd + bwd."""
    from oracle import configs as C
    cfg = C.tiny("pretrain", B=2, T=8, img=224, L=40, d=768, depth=2, heads=12, bert_layers=4, fusion_layer=2,
                 vocab=30522, num_entities=1000, seed=21)
    cfg["bert"]["max_position_embeddings"] = 512
    spec, sd, batch = helpers.make_inputs(cfg)
    sd_g, oout = helpers.oracle_run(cfg, sd, batch, requires_grad=True)
    sum(v for k, v in oout.items() if k.endswith("_loss") and v is not None).backward()
    model = build_cuda_model(cfg, sd)
    out = model(to_cuda(batch))
    assert out["_neg_video"].tolist() == oout["_neg_video"] and out["_neg_text"].tolist() == oout["_neg_text"]
    errs = {k: helpers.rel_err(out[k].detach().float().cpu(), oout[k].detach())
            for k in ("itc_loss", "itm_loss", "itm_scores", "mlm_loss", "mlm_scores", "mpm_loss", "mpm_logits", "mpm_labels")}
    errs["video_embeds"] = helpers.rel_err(out["_video_embeds"].cpu(), oout["_video_embeds"].detach())
    print("full-width", {k: f"{v:.2e}" for k, v in errs.items()})
    for k, v in errs.items():
        # mpm_labels: 1000-way teacher softmax at temperature ~0.07 behind 2 fp16 blocks of width 768 -- the label error
        # is the feature error (video_embeds, ~7e-4) amplified by the logit scale, hence its own bound
        tol = EMB_TOL if k.endswith("embeds") else (3e-3 if k == "mpm_labels" else FWD_TOL)
        assert v < tol, (k, v)
    sum(v for k, v in out.items() if k.endswith("_loss") and v is not None).backward()
    gmax = max(float(v.grad.abs().max()) for v in sd_g.values() if v.grad is not None)
    bad, errs_g = [], []
    for n, p in model.named_parameters():
        ref = sd_g[n].grad if n in sd_g else None
        if n.startswith("prompter.") or ref is None or float(ref.abs().max()) < 1e-6 * gmax:
            continue
        assert torch.isfinite(p.grad).all(), n
        e = helpers.rel_err(p.grad.cpu(), ref)
        errs_g.append((e, n))
        # mpm_head: d(loss)/d(logits) = softmax(logits) - soft_labels is a difference of two near-uniform 1000-way
        # distributions at synthetic init, so the ~6e-4 fp16 error of the teacher's labels is amplified ~50x there
        if e > (1e-1 if n.startswith("mpm_head.") else 3e-2):
            bad.append((n, e))
    print("full-width worst grads", [(n, f"{e:.2e}") for e, n in sorted(errs_g, reverse=True)[:6]])
    assert not bad, bad[:8]


def test_prefetch_loader_matches_plain_copies():
    """alpro_b200.prefetch.PrefetchLoader (reference: src/datasets/dataloader.py:80-157): same batches, same order,
    tensors on the device, optional normalisation applied to the visual keys only."""
    from alpro_b200.prefetch import PrefetchLoader
    g = torch.Generator().manual_seed(5)
    batches = [{"visual_inputs": torch.randint(0, 256, (2, 2, 3, 8, 8), dtype=torch.uint8, generator=g).pin_memory(),
                "text_input_ids": torch.randint(0, 100, (2, 6), generator=g).pin_memory(), "type": "video"}
               for _ in range(4)]
    # a batch lives in one of two staging sets and is overwritten two steps later: consume inside the loop
    n = 0
    for b, h in zip(PrefetchLoader(batches), batches):
        assert b["type"] == "video" and b["visual_inputs"].is_cuda and b["visual_inputs"].dtype == torch.uint8
        assert torch.equal(b["visual_inputs"].cpu(), h["visual_inputs"]) and torch.equal(b["text_input_ids"].cpu(), h["text_input_ids"])
        n += 1
    assert n == 4
    norm = lambda x: (x - 127.5) / 50.0
    n = 0
    for (task, b), h in zip(PrefetchLoader([("taskA", b) for b in batches], img_normalize=norm), batches):
        assert task == "taskA" and b["visual_inputs"].dtype == torch.float32
        assert torch.allclose(b["visual_inputs"].cpu(), norm(h["visual_inputs"].float()))
        assert torch.equal(b["text_input_ids"].cpu(), h["text_input_ids"])
        n += 1
    assert n == 4


@pytest.mark.parametrize("train", [False, True])
def test_fused_temporal_fc_matches_two_linears(train, monkeypatch):
    """ALPRO_FUSE_TFC=1 runs temporal_attn.proj -> DropPath -> temporal_fc (vit.py:157-161) as ONE GEMM with the composed
    weight W_fc W_proj (second, unscaled bias b_fc; gradients un-composed by two d x d GEMMs). Same losses and the same
    parameter gradients as the two-Linear path, in eval mode and with the DropPath draws of train mode replayed."""
    cfg = configs.GOLDEN["tiny_retrieval"]
    spec, sd, batch = helpers.make_inputs(cfg)
    results = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("ALPRO_FUSE_TFC", flag)
        model = build_cuda_model(cfg, sd)
        if train:
            model.train()
            torch.manual_seed(1234)          # same DropPath / dropout draws in both runs
            torch.cuda.manual_seed_all(1234)
            if hasattr(model.engine, "seed"):
                model.engine.seed = 77
        out = model(to_cuda(batch))
        loss = out["itc_loss"] + out["itm_loss"]
        loss.backward()
        torch.cuda.synchronize()
        results[flag] = ({k: out[k].detach().float().cpu() for k in ("itc_loss", "itm_loss", "itm_scores")},
                         {n: p.grad.detach().float().cpu() for n, p in model.named_parameters() if p.grad is not None})
    o0, g0 = results["0"]
    o1, g1 = results["1"]
    for k in o0:
        assert helpers.rel_err(o1[k], o0[k]) < FWD_TOL, k
    assert set(g0) == set(g1)
    # the composition only touches the visual encoder; gradients that are pure rounding noise in both runs (e.g. the
    # attention key biases, whose exact gradient is zero) are skipped by a floor relative to the largest gradient
    floor = 1e-4 * max(float(t.double().norm()) for t in g0.values())
    dev = []
    for n in g0:
        ref = float(g0[n].double().norm())
        if ref < floor or "visual_encoder" not in n:
            continue
        dev.append((float((g1[n].double() - g0[n].double()).norm()) / ref, n))
    dev.sort(reverse=True)
    assert len(dev) > 20
    print("fused temporal_fc: largest gradient deviations", [(n, f"{e:.2e}") for e, n in dev[:6]])
    print("losses", {k: (float(o0[k].flatten()[0]), float(o1[k].flatten()[0])) for k in o0})
    assert dev[0][0] < GRAD_TOL, dev[:6]
