"""GPU unit tests: every C-ABI kernel against a plain PyTorch fp32 statement of the same op (on the 16-bit-rounded
inputs the kernel actually sees). Run on the B200 box:  pytest tests -m gpu"""
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    from alpro_b200 import ops
    return ops


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def g(seed=0):
    return torch.Generator(device=DEV).manual_seed(seed)


# ------------------------------------------------------------------------------------------------------------ GEMM
GEMM_CASES = [
    # M, N, K, a_layout, b_layout, dtype, extras
    (128, 256, 64, 0, 0, torch.float16, {}),
    (1024, 768, 768, 0, 0, torch.float16, {}),
    (200, 296, 104, 0, 0, torch.float16, {}),
    (512, 512, 256, 0, 0, torch.bfloat16, {}),
    (640, 3072, 768, 0, 0, torch.float16, {"bias": 1, "act": 1, "out16b": 1}),
    (3 * 17, 768, 768, 0, 0, torch.float16, {"bias": 1, "resid": 1, "skip": 17, "out32": 1}),
    (1024, 768, 2304, 0, 1, torch.float16, {}),
    (640, 768, 3072, 0, 1, torch.float16, {"act": 2}),
    (200, 296, 104, 0, 1, torch.float16, {}),
    (2304, 768, 4096, 1, 1, torch.float16, {"split": -1}),
    (256, 512, 1000, 1, 1, torch.float16, {"out32": 1}),
    (256, 512, 1000, 1, 0, torch.float16, {"out32": 1}),
    (320, 1001, 192, 0, 0, torch.float16, {"bias": 1, "out32": 1}),          # unaligned N (vocab-style)
    (640, 192, 768, 0, 0, torch.float16, {"bias": 1, "act": 1, "out32": 1, "out16b": 1}),  # generic mode
    # second (never row-scaled) bias of the residual epilogue: TMA epilogue (N % 32 == 0) and the generic one
    (3 * 17 + 300, 768, 768, 0, 0, torch.float16, {"bias": 1, "bias2": 1, "resid": 1, "skip": 17, "out32": 1, "rs": 1}),
    (200, 296, 104, 0, 0, torch.float16, {"bias": 1, "bias2": 1, "resid": 1, "skip": 7, "out32": 1}),
]


# The GEMM has three implementations behind one entry point: the CTA-pair kernel with the TMA-store epilogue (default,
# gemm_tc3.cu), the CTA-pair kernel with the coalesced-store epilogue (ALPRO_GEMM_TMA_EPI=0, gemm_tc2.cu; also the
# generic / unaligned modes of the default path) and the single-CTA kernel (ALPRO_GEMM_2CTA=0, gemm_tc.cu).
@pytest.fixture(params=["tma_epilogue", "coalesced_epilogue", "single_cta"])
def gemm_impl(request, monkeypatch):
    monkeypatch.setenv("ALPRO_GEMM_TMA_EPI", "0" if request.param == "coalesced_epilogue" else "1")
    monkeypatch.setenv("ALPRO_GEMM_2CTA", "0" if request.param == "single_cta" else "1")
    return request.param


@pytest.mark.parametrize("case", GEMM_CASES, ids=[f"g{i}" for i in range(len(GEMM_CASES))])
def test_gemm16(case, gemm_impl):
    ops = _ops()
    M, N, K, al, bl, dt, ex = case
    gen = g(1)

    def mk(shape):
        ld = (shape[1] + 7) // 8 * 8
        buf = torch.zeros(shape[0], ld, device=DEV, dtype=dt)
        buf[:, :shape[1]] = (torch.randn(shape, device=DEV, generator=gen) * 0.5).to(dt)
        return buf[:, :shape[1]]

    a = mk((M, K) if al == 0 else (K, M))
    b = mk((N, K) if bl == 0 else (K, N))
    A = a.float() if al == 0 else a.float().t()
    Bm = b.float() if bl == 0 else b.float().t()
    ref = A @ Bm.t()
    kw = {}
    if ex.get("rs"):    # DropPath-style row factors on the accumulator and on the first bias
        kw["row_scale"] = (torch.rand(M, device=DEV, generator=gen) > 0.3).float() / 0.7
        ref = ref * kw["row_scale"][:, None]
    if ex.get("bias"):
        kw["bias"] = torch.randn(N, device=DEV, generator=gen)
        ref = ref + (kw["bias"] * kw["row_scale"][:, None] if ex.get("rs") else kw["bias"])
    if ex.get("bias2"):
        kw["bias2"] = torch.randn(N, device=DEV, generator=gen)
        ref = ref + kw["bias2"]
    pre = ref.clone()
    act = ex.get("act", 0)
    if act == 1:
        ref = F.gelu(ref)
    elif act == 2:   # aux holds the activation derivative saved by the forward epilogue
        aux = torch.randn(M, N, device=DEV, generator=gen).to(torch.float16)
        ref = ref * aux.float()
        kw["aux"] = aux
    kw["act"] = act
    if ex.get("resid"):
        resid = torch.randn(M, N, device=DEV, generator=gen)
        kw["resid"] = resid
        out = ref + resid
        sk = ex.get("skip", 0)
        if sk:
            rows = torch.arange(M, device=DEV) % sk == 0
            out[rows] = resid[rows]
            kw["skip_period"] = sk
        ref = out
    split = ex.get("split", 0)
    use32 = bool(ex.get("out32")) or split != 0
    if use32:
        out = torch.zeros(M * N + 4, device=DEV)[1:1 + M * N].view(M, N) if N % 4 else torch.zeros(M, N, device=DEV)
        kw["out32"] = out
    else:
        out = torch.zeros(M, (N + 7) // 8 * 8, device=DEV, dtype=dt)[:, :N]
        kw["out16"] = out
    out16b = None
    if ex.get("out16b"):
        out16b = torch.zeros(M, N, device=DEV, dtype=dt)
        kw["out16b"] = out16b
    ops.gemm16(a, b, a_layout=al, b_layout=bl, split_k=split, **kw)
    torch.cuda.synchronize()
    tol = 3e-5 * math.sqrt(K) if use32 else (2e-3 if dt == torch.float16 else 1.2e-2)
    assert rel(out.float(), ref) < tol
    if out16b is not None:   # GELU epilogue saves gelu'(pre)
        u = pre.clone().requires_grad_(True)
        F.gelu(u).sum().backward()
        assert rel(out16b.float(), u.grad) < 2e-3


# ------------------------------------------------------------------------------------------------------------ LN
@pytest.mark.parametrize("ln_path", ["bulk", "async", "reg"])
@pytest.mark.parametrize("d,eps,M", [(768, 1e-6, 333), (192, 1e-12, 333), (768, 1e-6, 4099), (256, 1e-6, 21000),
                                     (768, 1e-6, 9001)])
def test_layernorm_fwd_bwd(d, eps, M, ln_path, monkeypatch):
    """M >= 4096 and d = 768 takes the bulk-copy backward (layernorm_bulk.cu); ALPRO_LN_BWD_BULK=0 (or another d) the
    cp.async double-buffered one (one 12-warp block per SM); smaller M — or ALPRO_LN_BWD_ASYNC=0 — the register one."""
    monkeypatch.setenv("ALPRO_LN_BWD_BULK", "1" if ln_path == "bulk" else "0")
    monkeypatch.setenv("ALPRO_LN_FWD_BULK", "1" if ln_path == "bulk" else "0")
    monkeypatch.setenv("ALPRO_LN_BWD_ASYNC", "0" if ln_path == "reg" else "1")
    ops = _ops()
    gen = g(2)
    x = torch.randn(M, d, device=DEV, generator=gen) * 2 + 0.3
    gamma = 1 + 0.1 * torch.randn(d, device=DEV, generator=gen)
    beta = 0.1 * torch.randn(d, device=DEV, generator=gen)
    o32 = torch.empty(M, d, device=DEV)
    o16 = torch.empty(M, d, device=DEV, dtype=torch.float16)
    st = torch.empty(2, M, device=DEV)
    ops.layernorm_fwd(x, gamma, beta, eps, out32=o32, out16=o16, mean=st[0], rstd=st[1])
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (d,), gr, br, eps)
    assert rel(o32, ref.detach()) < 1e-5
    assert rel(o16.float(), ref.detach()) < 1e-3
    for kind in (torch.float32, torch.float16):
        dy = torch.randn(M, d, device=DEV, generator=gen)
        dyk = dy.to(kind)
        base = torch.randn(M, d, device=DEV, generator=gen)
        dx = base.clone()
        dx16 = torch.empty(M, d, device=DEV, dtype=torch.float16)
        dg = torch.zeros(d, device=DEV)
        db = torch.zeros(d, device=DEV)
        cs = torch.zeros(d, device=DEV)
        ops.layernorm_bwd(dyk, x, st[0], st[1], gamma, dx, 1, dx16=dx16, zero_period=7, dgamma=dg, dbeta=db,
                          param_scale=0.5, colsum=cs, colsum_zero_period=5)
        for t in (xr, gr, br):
            t.grad = None
        ref.backward(dyk.float(), retain_graph=True)
        assert rel(dx, base + xr.grad) < 2e-5
        assert rel(dg, 0.5 * gr.grad) < 1e-4 and rel(db, 0.5 * br.grad) < 1e-4
        keep5 = (torch.arange(M, device=DEV) % 5 != 0).float()[:, None]
        assert rel(cs, 0.5 * ((base + xr.grad) * keep5).sum(0)) < 1e-4
        want16 = (base + xr.grad).clone()
        want16[torch.arange(M, device=DEV) % 7 == 0] = 0
        assert rel(dx16.float(), want16) < 1e-3


@pytest.mark.parametrize("kind", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("accumulate", [0, 1])
def test_layernorm_bwd_bulk_hooks(kind, accumulate):
    """Bulk-copy LayerNorm backward (d = 768, M >= 4096) with every hook the TimeSformer backward uses: 16-bit / fp32
    upstream gradient, overwrite / accumulate, stochastic-depth row scales on the 16-bit copy and on the column sums,
    cls-row periods, bf16 output; checked against autograd."""
    ops = _ops()
    M, d, eps = 7 * 1571 + 3, 768, 1e-6
    gen = g(21)
    x = torch.randn(M, d, device=DEV, generator=gen) * 1.5 + 0.7
    gamma = 1 + 0.1 * torch.randn(d, device=DEV, generator=gen)
    beta = 0.1 * torch.randn(d, device=DEV, generator=gen)
    st = torch.empty(2, M, device=DEV)
    o16 = torch.empty(M, d, device=DEV, dtype=torch.float16)
    ops.layernorm_fwd(x, gamma, beta, eps, out16=o16, mean=st[0], rstd=st[1])
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (d,), gr, br, eps)
    dy = torch.randn(M, d, device=DEV, generator=gen).to(kind)
    ref.backward(dy.float())
    base = torch.randn(M, d, device=DEV, generator=gen)
    dx = base.clone()
    out_dt = torch.bfloat16 if kind == torch.bfloat16 else torch.float16
    dx16 = torch.empty(M, d, device=DEV, dtype=out_dt)
    rs16 = torch.rand(M, device=DEV, generator=gen) + 0.5
    rsc = torch.rand(M, device=DEV, generator=gen) + 0.5
    dg, db, cs = (torch.zeros(d, device=DEV) for _ in range(3))
    ops.layernorm_bwd(dy, x, st[0], st[1], gamma, dx, accumulate, dx16=dx16, zero_period=1571, dgamma=dg, dbeta=db,
                      param_scale=0.25, colsum=cs, colsum_zero_period=1571, dx16_row_scale=rs16, colsum_row_scale=rsc)
    want = xr.grad + (base if accumulate else 0)
    assert rel(dx, want) < 2e-5
    assert rel(dg, 0.25 * gr.grad) < 1e-4 and rel(db, 0.25 * br.grad) < 1e-4
    keep = (torch.arange(M, device=DEV) % 1571 != 0).float()[:, None]
    assert rel(cs, 0.25 * (want * keep * rsc[:, None]).sum(0)) < 1e-4
    assert rel(dx16.float(), want * keep * rs16[:, None]) < (1e-3 if out_dt == torch.float16 else 6e-3)
    # column sums default to the dx16 row scale when no separate scale is given
    cs2 = torch.zeros(d, device=DEV)
    dx2 = base.clone()
    ops.layernorm_bwd(dy, x, st[0], st[1], gamma, dx2, accumulate, dx16=dx16, colsum=cs2, dx16_row_scale=rs16)
    assert rel(cs2, (want * rs16[:, None]).sum(0)) < 1e-4


def test_colsum_cast():
    ops = _ops()
    x = torch.randn(1000, 770, device=DEV, generator=g(3))
    x16 = torch.empty(1000, 776, device=DEV, dtype=torch.float16)[:, :770]
    x16.copy_(x)
    out = torch.zeros(770, device=DEV)
    ops.colsum(x16, out, 0.25, zero_period=9)
    keep = (torch.arange(1000, device=DEV) % 9 != 0).float()[:, None]
    assert rel(out, 0.25 * (x16.float() * keep).sum(0)) < 1e-5
    out32 = torch.zeros(770, device=DEV)
    ops.colsum(x, out32, 2.0)
    assert rel(out32, 2.0 * x.sum(0)) < 1e-5
    src = torch.randn(100003, device=DEV, generator=g(4))
    dst = torch.empty(100003, device=DEV, dtype=torch.float16)
    ops.cast16(src, dst)
    assert torch.equal(dst, src.to(torch.float16))


# ------------------------------------------------------------------------------------------------------------ ViT glue
def test_patchify_embed_pool():
    ops = _ops()
    B, T, H, W, P, d = 2, 4, 48, 32, 16, 192
    gen = g(5)
    frames = torch.randn(B, T, 3, H, W, device=DEV, generator=gen)
    gw, gh = W // P, H // P
    N = gw * gh
    Sc = 1 + N * T
    out = torch.empty(B * Sc, 3 * P * P, device=DEV, dtype=torch.float16)
    ops.patchify(frames, out, P)
    # reference: unfold
    pt = frames.reshape(B * T, 3, gh, P, gw, P).permute(0, 2, 4, 1, 3, 5).reshape(B, T, N, 3 * P * P)
    want = torch.zeros(B, Sc, 3 * P * P, device=DEV)
    want[:, 1:] = pt.permute(0, 2, 1, 3).reshape(B, N * T, -1)
    assert torch.equal(out.view(B, Sc, -1), want.to(torch.float16))           # index work: bit-exact
    # raw uint8 frames with the input normalisation fused (ImageNorm)
    u8 = torch.randint(0, 256, (B, T, 3, H, W), device=DEV, generator=gen, dtype=torch.uint8)
    mean, std = (0.48, 0.45, 0.40), (0.27, 0.26, 0.28)
    out8 = torch.empty_like(out)
    ops.patchify_u8(u8, out8, P, mean, std)
    nf = (u8.float() / 255.0 - torch.tensor(mean, device=DEV).view(1, 1, 3, 1, 1)) / torch.tensor(std, device=DEV).view(1, 1, 3, 1, 1)
    ref8 = torch.empty_like(out)
    ops.patchify(nf.contiguous(), ref8, P)
    assert rel(out8.float(), ref8.float()) < 2e-3
    proj = torch.randn(B * Sc, d, device=DEV, generator=gen)
    cls, pos, tim = (torch.randn(s, device=DEV, generator=gen) for s in ((d,), (N + 1, d), (T, d)))
    x = torch.empty(B * Sc, d, device=DEV)
    ops.vit_embed_fwd(proj, cls, pos, tim, x, B, N, T, d)
    wx = proj.view(B, Sc, d).clone()
    wx[:, 0] = cls + pos[0]
    wx[:, 1:] = (proj.view(B, Sc, d)[:, 1:].view(B, N, T, d) + pos[1:].view(1, N, 1, d) + tim.view(1, 1, T, d)).view(B, N * T, d)
    assert torch.equal(x.view(B, Sc, d), wx) or rel(x.view(B, Sc, d), wx) < 1e-6
    dx = torch.randn(B * Sc, d, device=DEV, generator=gen)
    dcls, dpos, dtim = torch.zeros(d, device=DEV), torch.zeros(N + 1, d, device=DEV), torch.zeros(T, d, device=DEV)
    ops.vit_embed_bwd(dx, dcls, dpos, dtim, B, N, T, d, 0.5)
    dxv = dx.view(B, Sc, d)
    assert rel(dcls, 0.5 * dxv[:, 0].sum(0)) < 1e-5
    assert rel(dpos[1:], 0.5 * dxv[:, 1:].view(B, N, T, d).sum((0, 2))) < 1e-5 and rel(dpos[0], dcls) < 1e-6
    assert rel(dtim, 0.5 * dxv[:, 1:].view(B, N, T, d).sum((0, 1))) < 1e-5
    ve = torch.empty(B, 1 + N, d, device=DEV)
    ops.temporal_pool_fwd(x, ve, B, N, T, d)
    wv = torch.cat([x.view(B, Sc, d)[:, :1], x.view(B, Sc, d)[:, 1:].view(B, N, T, d).mean(2)], 1)
    assert rel(ve, wv) < 1e-6
    dve = torch.randn(B, 1 + N, d, device=DEV, generator=gen)
    dxn = torch.empty(B * Sc, d, device=DEV)
    ops.temporal_pool_bwd(dve, dxn, B, N, T, d)
    wd = torch.cat([dve[:, :1], (dve[:, 1:, None, :] / T).expand(B, N, T, d).reshape(B, N * T, d)], 1)
    assert rel(dxn.view(B, Sc, d), wd) < 1e-6


def test_bert_embed_and_fusion_gather():
    ops = _ops()
    B, L, h, V = 3, 8, 192, 500
    gen = g(6)
    ids = torch.randint(0, V, (B, L), device=DEV, generator=gen)
    word, pos, typ = (torch.randn(s, device=DEV, generator=gen) for s in ((V, h), (64, h), (2, h)))
    out = torch.empty(B * L, h, device=DEV)
    ops.bert_embed_gather(ids, word, pos, typ, out, L, h)
    want = word[ids] + typ[0] + pos[:L][None]
    assert rel(out.view(B, L, h), want) < 1e-6
    de = torch.randn(B * L, h, device=DEV, generator=gen)
    dw, dp, dt = torch.zeros_like(word), torch.zeros_like(pos), torch.zeros(h, device=DEV)
    ops.bert_embed_scatter(ids, de, dw, dp, dt, L, h, 0.5)
    ww = torch.zeros_like(word).index_add_(0, ids.view(-1), 0.5 * de)
    assert rel(dw, ww) < 1e-5 and rel(dp[:L], 0.5 * de.view(B, L, h).sum(0)) < 1e-5 and rel(dt, 0.5 * de.sum(0)) < 1e-5
    # fusion gather
    Nv, S = 5, 7
    text = torch.randn(2 * B, L, h, device=DEV, generator=gen)
    video = torch.randn(B, Nv, h, device=DEV, generator=gen)
    tmask = (torch.rand(2 * B, L, device=DEV, generator=gen) > 0.3).long()
    ti = torch.randint(0, 2 * B, (S,), device=DEV, generator=gen).int()
    vi = torch.randint(0, B, (S,), device=DEV, generator=gen).int()
    o32 = torch.empty(S * (L + Nv), h, device=DEV)
    o16 = torch.empty(S * (L + Nv), h, device=DEV, dtype=torch.float16)
    am = torch.empty(S, L + Nv, device=DEV)
    ops.fusion_gather_fwd(text, video, tmask, ti, vi, o32, o16, am, S, L, Nv, h)
    want = torch.cat([text[ti.long()], video[vi.long()]], 1)
    assert torch.equal(o32.view(S, L + Nv, h), want) and torch.equal(o16.view(S, L + Nv, h), want.half())
    wm = torch.cat([(1.0 - tmask[ti.long()].float()) * -10000.0, torch.zeros(S, Nv, device=DEV)], 1)
    assert torch.equal(am, wm)
    dout = torch.randn(S * (L + Nv), h, device=DEV, generator=gen)
    dtx, dvd = torch.zeros_like(text), torch.zeros_like(video)
    ops.fusion_gather_bwd(dout, ti, vi, dtx, dvd, S, L, Nv, h)
    dv = dout.view(S, L + Nv, h)
    assert rel(dtx, torch.zeros_like(text).index_add_(0, ti.long(), dv[:, :L])) < 1e-5
    assert rel(dvd, torch.zeros_like(video).index_add_(0, vi.long(), dv[:, L:])) < 1e-5


# ------------------------------------------------------------------------------------------------------------ attention
def _attn_ref(q, k, v, scale, mask=None):
    s = (q @ k.transpose(-1, -2)) * scale
    if mask is not None:
        s = s + mask
    return torch.softmax(s, -1) @ v


@pytest.mark.parametrize("T,N,impl", [(2, 5, "cuda_core"), (4, 5, "cuda_core"), (8, 5, "cuda_core"), (8, 5, "tcgen05"),
                                      (8, 16, "tcgen05"), (8, 196, "tcgen05")])
def test_temporal_attention(T, N, impl, monkeypatch):
    """T = 8 has two implementations (the library reads ALPRO_TATTN_TC on every call): CUDA-core registers/shuffles
    and the tcgen05 block-diagonal kernels (16 units per 128-row tile; N = 5 and 196 end in partial tiles)."""
    monkeypatch.setenv("ALPRO_TATTN_TC", "1" if impl == "tcgen05" else "0")
    ops = _ops()
    B, heads = 2, 3
    d = heads * 64
    Sc = 1 + N * T
    gen = g(7)
    qkv = (torch.randn(B * Sc, 3 * d, device=DEV, generator=gen)).half()
    out = torch.full((B * Sc, d), 7.0, device=DEV, dtype=torch.float16)
    ops.temporal_attn_fwd(qkv, out, B, N, T, heads, 0.125)
    x = qkv.float().view(B, Sc, 3, heads, 64)[:, 1:].reshape(B, N, T, 3, heads, 64).permute(3, 0, 1, 4, 2, 5)
    x = x.detach().requires_grad_(True)
    ref = _attn_ref(x[0], x[1], x[2], 0.125)                                  # [B,N,heads,T,64]
    refo = ref.permute(0, 1, 3, 2, 4).reshape(B, N * T, d)
    got = out.float().view(B, Sc, d)
    assert rel(got[:, 1:], refo.detach()) < 2e-3 and float(got[:, 0].abs().max()) == 0.0
    do = torch.randn(B * Sc, d, device=DEV, generator=gen).half()
    dqkv = torch.full((B * Sc, 3 * d), 7.0, device=DEV, dtype=torch.float16)
    ops.temporal_attn_bwd(qkv, do, dqkv, B, N, T, heads, 0.125)
    gq = do.float().view(B, Sc, d)[:, 1:].reshape(B, N, T, heads, 64).permute(0, 1, 3, 2, 4)
    ref.backward(gq)
    want = x.grad.permute(1, 2, 4, 0, 3, 5).reshape(B, N * T, 3 * d)
    gotd = dqkv.float().view(B, Sc, 3 * d)
    assert rel(gotd[:, 1:], want) < 3e-3 and float(gotd[:, 0].abs().max()) == 0.0


# forward: mma.sync (default) / tcgen05 (ALPRO_ATTN_TC=1); backward: mma.sync / tcgen05 (default for 96 <= S <= 240)
# the tcgen05 kernels load their operand tiles by tensor-map (TMA) boxes (default) or by the per-thread cp.async gather
# (ALPRO_ATTN_TMA=0)
# ALPRO_ATTN_PS=0: the per-unit tcgen05 forward instead of the persistent one (sattn_ps.cu, default for the strided
# two-tile layout without mask / dropout)
_ATTN_IMPLS = ["mma_sync", "tcgen05", "tcgen05_perunit", "tcgen05_bwd", "tcgen05_gather", "tcgen05_bwd_gather"]


@pytest.fixture(params=_ATTN_IMPLS)
def attn_impl(request, monkeypatch):
    """Sequence attention has mma.sync and tcgen05 implementations; the library reads ALPRO_ATTN_TC (forward),
    ALPRO_ATTN_BWD_TC (backward) and ALPRO_ATTN_TMA (operand loads of the tcgen05 kernels) on every call."""
    monkeypatch.setenv("ALPRO_ATTN_TC", "1" if request.param.startswith("tcgen05") and "bwd" not in request.param else "0")
    monkeypatch.setenv("ALPRO_ATTN_BWD_TC", "1" if request.param.startswith("tcgen05_bwd") else "0")
    monkeypatch.setenv("ALPRO_ATTN_TMA", "0" if request.param.endswith("gather") else "1")   # 1 = also for BERT rows
    monkeypatch.setenv("ALPRO_ATTN_PS", "0" if request.param == "tcgen05_perunit" else "1")
    return request.param


@pytest.mark.parametrize("S,dt", [(40, torch.float16), (237, torch.float16), (21, torch.bfloat16), (256, torch.float16),
                                  (197, torch.float16), (129, torch.bfloat16)])
def test_seq_attention_bert_layout(S, dt, attn_impl):
    ops = _ops()
    nseq, heads = 3, 3
    d = heads * 64
    gen = g(8)
    qkv = torch.randn(nseq * S, 3 * d, device=DEV, generator=gen).to(dt)
    keep = torch.rand(nseq, S, device=DEV, generator=gen) > 0.25
    keep[:, 0] = True
    mask = (1.0 - keep.float()) * -10000.0
    o = torch.empty(nseq * S, d, device=DEV, dtype=dt)
    lse = torch.empty(nseq, heads, S, device=DEV)
    scale = 1 / math.sqrt(64)
    ops.seq_attn_fwd(qkv, mask, o, None, lse, S, nseq, heads, 1, 1, S, scale)
    x = qkv.float().view(nseq, S, 3, heads, 64).permute(2, 0, 3, 1, 4).detach().requires_grad_(True)
    ref = _attn_ref(x[0], x[1], x[2], scale, mask[:, None, None, :])          # [nseq,heads,S,64]
    refo = ref.permute(0, 2, 1, 3).reshape(nseq * S, d)
    tol = 3e-3 if dt == torch.float16 else 2e-2
    assert rel(o.float(), refo.detach()) < tol
    sc = torch.einsum("nhid,nhjd->nhij", x[0].detach(), x[1].detach()) * scale + mask[:, None, None, :]
    assert float((lse * math.log(2.0) - torch.logsumexp(sc, -1)).abs().max()) < 2e-3      # lse is kept in base 2
    do = torch.randn(nseq * S, d, device=DEV, generator=gen).to(dt)
    dqkv = torch.empty(nseq * S, 3 * d, device=DEV, dtype=dt)
    ops.seq_attn_bwd(qkv, mask, lse, o, None, do, dqkv, None, S, nseq, heads, 1, 1, S, scale)
    ref.backward(do.float().view(nseq, S, heads, 64).permute(0, 2, 1, 3))
    want = x.grad.permute(1, 3, 0, 2, 4).reshape(nseq * S, 3 * d)
    assert rel(dqkv.float(), want) < (4e-3 if dt == torch.float16 else 3e-2)


@pytest.mark.parametrize("N,T", [(4, 2), (196, 2), (9, 4)])
def test_seq_attention_vit_layout(N, T, attn_impl):
    """'(b t) (h w)' spatial attention read from the canonical 'b (h w t)' rows, shared cls row, cls mean."""
    ops = _ops()
    B, heads = 2, 3
    d = heads * 64
    Sc = 1 + N * T
    S = 1 + N
    gen = g(9)
    qkv = torch.randn(B * Sc, 3 * d, device=DEV, generator=gen).half()
    o = torch.zeros(B * Sc, d, device=DEV, dtype=torch.float16)
    cls_o = torch.empty(B * T, d, device=DEV, dtype=torch.float16)
    lse = torch.empty(B * T, heads, S, device=DEV)
    ops.seq_attn_fwd(qkv, None, o, cls_o, lse, S, B * T, heads, T, T, Sc, 0.125)
    ops.cls_mean_fwd(cls_o, o, B, T, Sc, d)
    # reference on explicitly rearranged tensors
    x = qkv.float().view(B, Sc, 3 * d).detach().requires_grad_(True)
    cls = x[:, :1].unsqueeze(1).expand(B, T, 1, 3 * d)
    pat = x[:, 1:].view(B, N, T, 3 * d).permute(0, 2, 1, 3)
    xs = torch.cat([cls, pat], 2).reshape(B * T, S, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = _attn_ref(xs[0], xs[1], xs[2], 0.125).permute(0, 2, 1, 3).reshape(B, T, S, d)
    want = torch.cat([ref[:, :, 0].mean(1, keepdim=True), ref[:, :, 1:].permute(0, 2, 1, 3).reshape(B, N * T, d)], 1)
    assert rel(o.float().view(B, Sc, d), want.detach()) < 3e-3
    do = torch.randn(B * Sc, d, device=DEV, generator=gen).half()
    dqkv = torch.empty(B * Sc, 3 * d, device=DEV, dtype=torch.float16)
    scratch = torch.empty(B * T, 3 * d, device=DEV)
    ops.seq_attn_bwd(qkv, None, lse, o, cls_o, do, dqkv, scratch, S, B * T, heads, T, T, Sc, 0.125)
    want.backward(do.float().view(B, Sc, d))
    assert rel(dqkv.float().view(B, Sc, 3 * d), x.grad) < 4e-3


@pytest.mark.parametrize("B,T,N,heads,dt", [(2, 8, 196, 3, torch.float16), (1, 4, 160, 2, torch.float16),
                                            (1, 2, 223, 2, torch.bfloat16), (3, 8, 196, 12, torch.bfloat16),
                                            (1, 2, 129, 1, torch.float16)])
def test_seq_attention_persistent_forward(B, T, N, heads, dt, monkeypatch):
    """Persistent warp-specialised forward (sattn_ps.cu: operand ring, probabilities as a TMEM A operand, tensor-map
    stores) against the per-unit tcgen05 kernel and a torch fp32 reference: patch rows, per-frame cls outputs, base-2
    log-sum-exp in token order; the clip's cls row of o must stay untouched (cls_mean_fwd writes it later)."""
    ops = _ops()
    monkeypatch.setenv("ALPRO_ATTN_TC", "1")
    d = heads * 64
    Sc, S, nseq = 1 + N * T, 1 + N, B * T
    gen = g(31)
    qkv = torch.randn(B * Sc, 3 * d, device=DEV, generator=gen).to(dt)
    outs = {}
    for name, flag in (("unit", "0"), ("ps", "1")):
        monkeypatch.setenv("ALPRO_ATTN_PS", flag)
        o = torch.full((B * Sc, d), 7.0, device=DEV, dtype=dt)
        cls_o = torch.full((nseq, d), 7.0, device=DEV, dtype=dt)
        lse = torch.full((nseq, heads, S), 7.0, device=DEV)
        ops.seq_attn_fwd(qkv, None, o, cls_o, lse, S, nseq, heads, T, T, Sc, 0.125)
        torch.cuda.synchronize()
        outs[name] = (o.float().view(B, Sc, d), cls_o.float(), lse)
    x = qkv.float().view(B, Sc, 3 * d)
    cls = x[:, :1].unsqueeze(1).expand(B, T, 1, 3 * d)
    pat = x[:, 1:].view(B, N, T, 3 * d).permute(0, 2, 1, 3)
    xs = torch.cat([cls, pat], 2).reshape(B * T, S, 3, heads, 64).permute(2, 0, 3, 1, 4)
    sc = (xs[0] @ xs[1].transpose(-1, -2)) * 0.125
    ref = (torch.softmax(sc, -1) @ xs[2]).permute(0, 2, 1, 3).reshape(B, T, S, d)
    want = ref[:, :, 1:].permute(0, 2, 1, 3).reshape(B, N * T, d)
    tol = 3e-3 if dt == torch.float16 else 2e-2
    for name in ("unit", "ps"):
        o, c, lse = outs[name]
        assert rel(o[:, 1:], want) < tol and rel(c.view(B, T, d), ref[:, :, 0]) < tol
        assert bool((o[:, 0] == 7.0).all())
        assert float((lse * math.log(2.0) - torch.logsumexp(sc, -1)).abs().max()) < 2e-3
    # the two kernels round the same fp32 results: they may differ by one unit in the last place of a 16-bit output
    ulp = 2.0 ** -10 if dt == torch.float16 else 2.0 ** -7
    assert rel(outs["ps"][0][:, 1:], outs["unit"][0][:, 1:]) < 2 * ulp
    assert float((outs["ps"][2] - outs["unit"][2]).abs().max()) < 1e-4


# ------------------------------------------------------------------------------------------------------------ heads
@pytest.mark.parametrize("M,N,K,acc", [(1, 768, 768, False), (32, 1000, 1536, False), (5, 256, 768, True), (3, 130, 100, True)])
def test_small_linear_dx_long_contraction(M, N, K, acc):
    """dx = alpha * dy' W for long N: the block-split kernel (N >= 128, atomics into a zeroed / accumulated dx) with a
    ReLU gate, strided dx rows and both accumulate modes."""
    ops = _ops()
    gen = g(12)
    dy = torch.randn(M, N, device=DEV, generator=gen)
    yact = torch.randn(M, N, device=DEV, generator=gen)
    W = torch.randn(N, K, device=DEV, generator=gen)
    ld = K + 12
    dx = torch.randn(M, ld, device=DEV, generator=gen)
    before = dx.clone()
    ops.small_linear_bwd(dy, N, yact, None, 0, W, dx, ld, int(acc), None, None, 0, M, N, K, alpha=0.5)
    want = 0.5 * ((dy * (yact > 0)) @ W)
    if acc:
        want = want + before[:, :K]
    assert rel(dx[:, :K], want) < 1e-5
    assert torch.equal(dx[:, K:], before[:, K:])          # the padding columns of the strided rows are untouched


def test_cast_multi_matches_single_casts():
    """OperandCache.refresh: one batched launch == the per-tensor casts (ragged sizes, both formats)."""
    from alpro_b200.engine import OperandCache
    for dt in (torch.float16, torch.bfloat16):
        gen = g(13)
        ps = {f"w{i}": torch.nn.Parameter(torch.randn(shape, device=DEV, generator=gen))
              for i, shape in enumerate([(768, 768), (3, 5), (16384 * 2 + 4,), (1000, 7), (64, 192)])}
        W = OperandCache(dt)
        first = {n: W.get(n, p).clone() for n, p in ps.items()}
        for n, p in ps.items():
            assert torch.equal(first[n], p.detach().to(dt))
        with torch.no_grad():
            for p in ps.values():
                v0 = p._version
                p.data.mul_(1.5).add_(0.25)                  # .data update: the version counter does not move
                assert p._version == v0
        W.refresh()                                          # ONE alpro_cast_f32_to_16_multi launch
        calls = _ops()._L.calls
        for n, p in ps.items():
            assert torch.equal(W.get(n, p), p.detach().to(dt)), n
        assert _ops()._L.calls == calls                      # every get() was a hit: nothing re-cast lazily


def test_small_linear_l2norm():
    ops = _ops()
    gen = g(10)
    M, N, K, ldx = 6, 50, 192, 192 * 5
    xbig = torch.randn(M, 5, K, device=DEV, generator=gen)
    W, b = torch.randn(N, K, device=DEV, generator=gen), torch.randn(N, device=DEV, generator=gen)
    temp = torch.tensor(0.07, device=DEV)
    y = torch.empty(M, N, device=DEV)
    ops.small_linear_fwd(xbig, ldx, W, b, y, M, N, K, 1.0, temp, 2, relu=True)
    x0 = xbig[:, 0].clone().requires_grad_(True)
    Wr, br = W.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.relu((x0 @ Wr.t()) / temp + br)
    # bias is added after alpha scaling in the kernel: y = alpha * xW^T + b
    assert rel(y, ref.detach()) < 1e-5
    dy = torch.randn(M, N, device=DEV, generator=gen)
    dx = torch.zeros(M, 5, K, device=DEV)
    dW, db = torch.empty(N, K, device=DEV), torch.empty(N, device=DEV)
    ops.small_linear_bwd(dy, N, y, xbig, ldx, W, dx, ldx, 1, dW, db, 0, M, N, K, alpha=1.0, alpha_dev=temp,
                         alpha_mode=2, dw_scale=0.5)
    ref.backward(dy)
    assert rel(dx[:, 0], x0.grad) < 1e-5 and float(dx[:, 1:].abs().max()) == 0.0
    assert rel(dW, 0.5 * Wr.grad) < 1e-5 and rel(db, 0.5 * br.grad) < 1e-5
    x = torch.randn(7, 256, device=DEV, generator=gen).requires_grad_(True)
    yv, nrm = torch.empty(7, 256, device=DEV), torch.empty(7, device=DEV)
    ops.l2norm_fwd(x.detach(), yv, nrm)
    r = F.normalize(x, dim=-1)
    assert rel(yv, r.detach()) < 1e-6
    dyv = torch.randn(7, 256, device=DEV, generator=gen)
    dxv = torch.empty(7, 256, device=DEV)
    ops.l2norm_bwd(dyv, yv, nrm, dxv)
    r.backward(dyv)
    assert rel(dxv, x.grad) < 1e-5


def test_softmax_ce_variants():
    ops = _ops()
    gen = g(11)
    R, C = 37, 1003
    logits = torch.randn(R, C, device=DEV, generator=gen) * 3
    hard = torch.randint(0, C, (R,), device=DEV, generator=gen)
    hard[::5] = -100
    st = ops.softmax_ce_fwd(logits, C, hard=hard, denom_mode=0)
    lr = logits.clone().requires_grad_(True)
    ref = F.cross_entropy(lr, hard, ignore_index=-100)
    assert rel(st.loss, ref.detach()) < 1e-5
    gup = torch.tensor(0.7, device=DEV)
    Cp = (C + 7) // 8 * 8
    d16 = torch.full((R, Cp), 5.0, device=DEV, dtype=torch.float16)
    d32 = torch.empty(R, C, device=DEV)
    ops.softmax_ce_bwd(logits, C, st, gup, 64.0, hard=hard, out16=d16, C_out=Cp)
    ops.softmax_ce_bwd(logits, C, st, gup, 64.0, hard=hard, out32=d32)
    ref.backward(gup)
    assert rel(d32, 64.0 * lr.grad) < 1e-5 and rel(d16[:, :C].float(), 64.0 * lr.grad) < 2e-3
    assert float(d16[:, C:].abs().max()) == 0.0
    # soft labels with ignored rows (MPM) and mean over all rows (VTC / VTM)
    soft = torch.softmax(torch.randn(R, C, device=DEV, generator=gen), 1)
    ign = (torch.rand(R, device=DEV, generator=gen) < 0.3)
    st2 = ops.softmax_ce_fwd(logits, C, soft=soft, row_ignore=ign.to(torch.uint8), denom_mode=0)
    lr2 = logits.clone().requires_grad_(True)
    ce = -(F.log_softmax(lr2, 1) * soft).sum(1)
    ce = torch.where(ign, torch.zeros_like(ce), ce)
    ref2 = ce.sum() / (R - ign.sum())
    assert rel(st2.loss, ref2.detach()) < 1e-5
    ops.softmax_ce_bwd(logits, C, st2, None, 1.0, soft=soft, out32=d32)
    ref2.backward()
    assert rel(d32, lr2.grad) < 1e-5
    st3 = ops.softmax_ce_fwd(logits, C, hard=hard.clamp_min(0), denom_mode=1)
    assert rel(st3.loss, F.cross_entropy(logits, hard.clamp_min(0))) < 1e-5


def test_misc_heads():
    ops = _ops()
    gen = g(12)
    b = 9
    sim = torch.randn(b, 3 * b, device=DEV, generator=gen)
    w = torch.empty(b, b, device=DEV)
    ops.neg_weights(sim, b, b, w)
    blk = sim[:, b:2 * b].clone()
    blk.fill_diagonal_(-float("inf"))
    assert rel(w, torch.softmax(blk, 1)) < 1e-5
    B, R, L, Np, h = 3, 14, 4, 9, 192
    fo = torch.randn(2 * B * R, h, device=DEV, generator=gen)
    pm = (torch.rand(B, Np, device=DEV, generator=gen) > 0.5).float()
    pm[:, 0] = 0
    pooled = torch.empty(B, h, device=DEV)
    ops.masked_mean_fwd(fo, R * h, L + 1, pm, B, Np, h, pooled)
    vis = fo.view(2 * B, R, h)[:B, L + 1:L + 1 + Np]
    inv = (1 - pm)[:, :, None]
    assert rel(pooled, (inv * vis).sum(1) / inv.sum(1)) < 1e-5
    dpool = torch.randn(B, h, device=DEV, generator=gen)
    dfo = torch.zeros_like(fo)
    ops.masked_mean_bwd(dpool, pm, B, Np, h, dfo, R * h, L + 1)
    want = torch.zeros(2 * B, R, h, device=DEV)
    want[:B, L + 1:L + 1 + Np] = inv * dpool[:, None, :] / inv.sum(1, keepdim=True)
    assert rel(dfo.view(2 * B, R, h), want) < 1e-5
    o16 = torch.empty(B * L, h, device=DEV, dtype=torch.float16)
    ops.take_rows_fwd(fo, R, B, B, L, h, out16=o16)
    assert torch.equal(o16.view(B, L, h), fo.view(2 * B, R, h)[B:, :L].half())
    d2 = torch.zeros_like(fo)
    dd = torch.randn(B * L, h, device=DEV, generator=gen)
    ops.take_rows_bwd(dd, R, B, B, L, h, d2)
    w2 = torch.zeros(2 * B, R, h, device=DEV)
    w2[B:, :L] = dd.view(B, L, h)
    assert torch.equal(d2.view(2 * B, R, h), w2)
    simp = torch.randn(6, 48, device=DEV, generator=gen)
    simp[2, 0] = 100.0
    soft, ign = torch.empty(6, 48, device=DEV), torch.empty(6, device=DEV, dtype=torch.uint8)
    ops.pseudo_labels(simp, soft, ign)
    assert rel(soft, torch.softmax(simp, 1)) < 1e-5
    assert torch.equal(ign.bool(), torch.max(simp, 1)[1] < 0.2)
    dy = torch.randn(50, 192, device=DEV, generator=gen)
    pre = torch.randn(50, 192, device=DEV, generator=gen).half()
    out = torch.empty(50, 192, device=DEV, dtype=torch.float16)
    ops.gelu_grad_mul(dy, pre, out)
    assert rel(out.float(), dy * pre.float()) < 2e-3
    t = torch.tensor(0.9, device=DEV)
    ops.clamp_scalar(t, 0.001, 0.5)
    assert float(t) == 0.5
