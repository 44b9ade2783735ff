"""Python -> C-ABI call layer. Each function takes torch CUDA tensors (used only as device-memory handles), checks
dtype/contiguity and forwards raw pointers + sizes + the current CUDA stream to libalpro_b200.so."""
import ctypes

import torch

from . import _lib
from ._lib import GemmEpilogue, check

F16, BF16 = 0, 1
KMAJOR, MNMAJOR = 0, 1
ACT_NONE, ACT_GELU, ACT_GELU_GRAD, ACT_RELU, ACT_RELU_GRAD = 0, 1, 2, 3, 4

_FMT = {torch.float16: F16, torch.bfloat16: BF16}
GEMM_PROFILE = None  # set to a list to record (M, N, K, start_event, end_event) per GEMM launch (bench.py roofline)


def _stream():
    return ctypes.c_void_p(_s())


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _fmt(t):
    try:
        return _FMT[t.dtype]
    except KeyError:
        raise TypeError(f"expected a float16/bfloat16 tensor, got {t.dtype}")


def _check_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("alpro_b200 kernels require CUDA tensors (no CPU fallback)")


def _row_major_2d(t, name):
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name}: expected a 2-D tensor with unit inner stride, got {tuple(t.shape)} / {t.stride()}")
    return t.stride(0)


def gemm16(a, b, *, a_layout=KMAJOR, b_layout=KMAJOR, bias=None, act=ACT_NONE, aux=None, resid=None, out32=None,
           out16=None, out16b=None, skip_period=0, split_k=0, alpha=1.0, row_scale=None, row_scale_bias=None,
           bias2=None):
    """acc[m,n] = sum_k A(m,k) B(n,k) on tcgen05 tensor cores, fused epilogue (see include/alpro_b200.h).

    a: [M,K] (K-major) or [K,M] (MN-major) 16-bit; b: [N,K] (K-major) or [K,N] (MN-major) 16-bit.
    """
    _check_cuda(a, b, bias, aux, resid, out32, out16, out16b)
    lda = _row_major_2d(a, "a")
    ldb = _row_major_2d(b, "b")
    if a_layout == KMAJOR:
        M, K = a.shape
    else:
        K, M = a.shape
    if b_layout == KMAJOR:
        N, Kb = b.shape
    else:
        Kb, N = b.shape
    if K != Kb:
        raise ValueError(f"gemm16: contraction mismatch {K} vs {Kb}")
    ep = GemmEpilogue()
    ep.bias = _ptr(bias)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N and bias.is_contiguous()
    ep.aux16 = _ptr(aux)
    if aux is not None:
        ep.ldaux = _row_major_2d(aux, "aux")
        ep.aux_fmt = _fmt(aux)
        assert tuple(aux.shape) == (M, N)
    ep.resid = _ptr(resid)
    if resid is not None:
        assert resid.dtype == torch.float32 and tuple(resid.shape) == (M, N)
        ep.ldresid = _row_major_2d(resid, "resid")
    ep.out32 = _ptr(out32)
    if out32 is not None:
        assert out32.dtype == torch.float32 and tuple(out32.shape) == (M, N)
        ep.ld32 = _row_major_2d(out32, "out32")
    ep.out16 = _ptr(out16)
    if out16 is not None:
        assert tuple(out16.shape) == (M, N)
        ep.ld16 = _row_major_2d(out16, "out16")
        ep.out16_fmt = _fmt(out16)
    ep.out16b = _ptr(out16b)
    if out16b is not None:
        assert tuple(out16b.shape) == (M, N)
        ep.ld16b = _row_major_2d(out16b, "out16b")
        ep.out16b_fmt = _fmt(out16b)
    ep.act = act
    ep.skip_period = skip_period
    ep.split_k = split_k
    ep.alpha = alpha
    ep.row_scale_acc = _ptr(row_scale)
    ep.row_scale_bias = _ptr(row_scale_bias)
    ep.bias2 = _ptr(bias2)
    if bias2 is not None:
        assert bias2.dtype == torch.float32 and bias2.numel() == N and bias2.is_contiguous()
    prof = GEMM_PROFILE
    if prof is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    check(_lib.counted.alpro_gemm16(_ptr(a), _ptr(b), M, N, K, lda, ldb, a_layout, b_layout, _fmt(a), _fmt(b),
                                    ctypes.byref(ep), _stream()), "alpro_gemm16")
    if prof is not None:
        e1.record()
        # algorithmic bytes of this launch: both 16-bit operands once + every epilogue stream at its real width
        nbytes = 2.0 * (M * K + N * K) + M * N * (4.0 * (out32 is not None) + 2.0 * (out16 is not None)
                                                   + 2.0 * (out16b is not None) + 4.0 * (resid is not None)
                                                   + 2.0 * (aux is not None))
        tag = ("A" + ("mn" if a_layout == MNMAJOR else "k") + " B" + ("mn" if b_layout == MNMAJOR else "k")
               + (" splitK" if split_k else "") + (" act%d" % act if act else "") + (" resid" if resid is not None else "")
               + (" o32" if out32 is not None else "") + (" o16" if out16 is not None else "")
               + (" o16b" if out16b is not None else "") + (" aux" if aux is not None else "")
               + (" rs" if row_scale is not None else ""))
        prof.append((M, N, K, e0, e1, nbytes, tag))


# ---------------------------------------------------------------------------------------------------------------------
# thin wrappers for the remaining entry points (argument order = include/alpro_b200.h)
# ---------------------------------------------------------------------------------------------------------------------
_KIND = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}
_L = _lib.counted


def _p(t):
    return t.data_ptr() if t is not None else None


# raw cudaStream_t of torch's current stream: the private C getters cost ~0.3 us, torch.cuda.current_stream() ~13 us
# (x 924 launches per step = 12 ms of host time, tools/host_profile.py)
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_cur_device = getattr(torch._C, "_cuda_getDevice", None)


def _s():
    if _raw_stream is not None and _cur_device is not None:
        return _raw_stream(_cur_device())
    return torch.cuda.current_stream().cuda_stream


def cast16(src, dst):
    assert src.dtype == torch.float32 and src.is_contiguous() and dst.is_contiguous() and src.numel() == dst.numel()
    check(_L.alpro_cast_f32_to_16(_p(src), _p(dst), src.numel(), _fmt(dst), _s()), "alpro_cast_f32_to_16")


CAST_CHUNK = 16384


def cast16_multi(table, num_chunks, fmt):
    check(_L.alpro_cast_f32_to_16_multi(_p(table), num_chunks, fmt, _s()), "alpro_cast_f32_to_16_multi")


def layernorm_fwd(x, gamma, beta, eps, out32=None, out16=None, mean=None, rstd=None, mul16=None):
    M, d = x.shape
    check(_L.alpro_layernorm_fwd(_p(x), x.stride(0), _p(gamma), _p(beta), eps, M, d, _p(out32),
                                 out32.stride(0) if out32 is not None else 0, _p(out16),
                                 out16.stride(0) if out16 is not None else 0,
                                 _fmt(out16) if out16 is not None else (_fmt(mul16) if mul16 is not None else 0),
                                 _p(mean), _p(rstd), _p(mul16), mul16.stride(0) if mul16 is not None else 0, _s()),
          "alpro_layernorm_fwd")


def layernorm_bwd(dy, x, mean, rstd, gamma, dx32, accumulate, dx16=None, zero_period=0, dgamma=None, dbeta=None,
                  param_scale=1.0, colsum=None, colsum_zero_period=0, dy_mul16=None, dx16_mul16=None,
                  dx16_row_scale=None, colsum_row_scale=None):
    M, d = x.shape
    check(_L.alpro_layernorm_bwd(_p(dy), _KIND[dy.dtype], dy.stride(0), _p(x), x.stride(0), _p(mean), _p(rstd),
                                 _p(gamma), M, d, _p(dx32), dx32.stride(0), int(accumulate), _p(dx16),
                                 dx16.stride(0) if dx16 is not None else 0,
                                 _fmt(dx16) if dx16 is not None else (_fmt(dy_mul16) if dy_mul16 is not None else 0),
                                 zero_period, _p(dgamma), _p(dbeta), param_scale, _p(colsum), colsum_zero_period,
                                 _p(dy_mul16), dy_mul16.stride(0) if dy_mul16 is not None else 0, _p(dx16_mul16),
                                 dx16_mul16.stride(0) if dx16_mul16 is not None else 0, _p(dx16_row_scale),
                                 _p(colsum_row_scale), _s()), "alpro_layernorm_bwd")


def colsum(x, out, alpha=1.0, zero_period=0):
    M, N = x.shape
    check(_L.alpro_colsum(_p(x), _KIND[x.dtype], x.stride(0), M, N, _p(out), alpha, zero_period, _s()), "alpro_colsum")


def patchify(frames, out16, P):
    B, T, C, H, W = frames.shape
    assert C == 3 and frames.is_contiguous() and frames.dtype == torch.float32
    check(_L.alpro_patchify(_p(frames), _p(out16), _fmt(out16), B, T, H, W, P, _s()), "alpro_patchify")


def patchify_u8(frames, out16, P, mean, std):
    B, T, C, H, W = frames.shape
    assert C == 3 and frames.is_contiguous() and frames.dtype == torch.uint8
    m = (ctypes.c_float * 3)(*[float(v) for v in mean])
    sd = (ctypes.c_float * 3)(*[float(v) for v in std])
    check(_L.alpro_patchify_u8(_p(frames), _p(out16), _fmt(out16), B, T, H, W, P, ctypes.cast(m, ctypes.c_void_p),
                               ctypes.cast(sd, ctypes.c_void_p), _s()), "alpro_patchify_u8")


def vit_embed_fwd(proj, cls, pos, tim, x, B, N, T, d):
    check(_L.alpro_vit_embed_fwd(_p(proj), _p(cls), _p(pos), _p(tim), _p(x), B, N, T, d, _s()), "alpro_vit_embed_fwd")


def vit_embed_bwd(dx, dcls, dpos, dtim, B, N, T, d, alpha):
    check(_L.alpro_vit_embed_bwd(_p(dx), _p(dcls), _p(dpos), _p(dtim), B, N, T, d, alpha, _s()), "alpro_vit_embed_bwd")


def temporal_pool_fwd(xn, out, B, N, T, d):
    check(_L.alpro_temporal_pool_fwd(_p(xn), _p(out), B, N, T, d, _s()), "alpro_temporal_pool_fwd")


def temporal_pool_bwd(dout, dxn, B, N, T, d, alpha=1.0):
    check(_L.alpro_temporal_pool_bwd(_p(dout), _p(dxn), B, N, T, d, alpha, _s()), "alpro_temporal_pool_bwd")


def bert_embed_gather(ids, word, pos, type_, out, L, h):
    check(_L.alpro_bert_embed_gather(_p(ids), _p(word), _p(pos), _p(type_), _p(out), ids.numel(), L, h, _s()),
          "alpro_bert_embed_gather")


def bert_embed_scatter(ids, de, dword, dpos, dtype_, L, h, alpha):
    check(_L.alpro_bert_embed_scatter(_p(ids), _p(de), _p(dword), _p(dpos), _p(dtype_), ids.numel(), L, h, alpha, _s()),
          "alpro_bert_embed_scatter")


def fusion_gather_fwd(text, video, tmask, ti, vi, out32, out16, add_mask, S, L, Nv, h):
    check(_L.alpro_fusion_gather_fwd(_p(text), _p(video), _p(tmask), _p(ti), _p(vi), _p(out32), _p(out16),
                                     _fmt(out16) if out16 is not None else 0, _p(add_mask), S, L, Nv, h, _s()),
          "alpro_fusion_gather_fwd")


def fusion_gather_bwd(dout, ti, vi, dtext, dvideo, S, L, Nv, h):
    check(_L.alpro_fusion_gather_bwd(_p(dout), _p(ti), _p(vi), _p(dtext), _p(dvideo), S, L, Nv, h, _s()),
          "alpro_fusion_gather_bwd")


def cls_mean_fwd(cls_t, o, B, T, clip_rows, d, frame_weight=None):
    check(_L.alpro_cls_mean_fwd(_p(cls_t), _p(o), o.stride(0), _fmt(o), B, T, clip_rows, d, _p(frame_weight), _s()),
          "alpro_cls_mean_fwd")


def dropout_mask(out16, p, seed):
    check(_L.alpro_dropout_mask(_p(out16), _fmt(out16), out16.numel(), p, seed & 0xffffffff, _s()), "alpro_dropout_mask")


def temporal_attn_fwd(qkv, out, B, N, T, heads, scale):
    check(_L.alpro_temporal_attn_fwd(_p(qkv), qkv.stride(0), _p(out), out.stride(0), B, N, T, heads, _fmt(qkv), scale,
                                     _s()), "alpro_temporal_attn_fwd")


def temporal_attn_bwd(qkv, dout, dqkv, B, N, T, heads, scale):
    check(_L.alpro_temporal_attn_bwd(_p(qkv), qkv.stride(0), _p(dout), dout.stride(0), _p(dqkv), dqkv.stride(0), B, N,
                                     T, heads, _fmt(qkv), scale, _s()), "alpro_temporal_attn_bwd")


def seq_attn_fwd(qkv, mask, o, cls_o, lse, S, nseq, heads, seq_div, stride, clip_rows, scale, drop_p=0.0, drop_seed=0):
    check(_L.alpro_seq_attn_fwd(_p(qkv), qkv.stride(0), _p(mask), _p(o), o.stride(0), _p(cls_o), _p(lse), S, nseq,
                                heads, _fmt(qkv), seq_div, stride, clip_rows, scale, drop_p, drop_seed & 0xffffffff,
                                _s()), "alpro_seq_attn_fwd")


def attn_dropout_mask(out, S, nseq, heads, drop_p, drop_seed):
    check(_L.alpro_attn_dropout_mask(_p(out), S, nseq, heads, drop_p, drop_seed & 0xffffffff, _s()),
          "alpro_attn_dropout_mask")


def seq_attn_bwd(qkv, mask, lse, o_fwd, cls_fwd, dout, dqkv, scratch, S, nseq, heads, seq_div, stride, clip_rows, scale,
                 cls_weight=None, drop_p=0.0, drop_seed=0):
    assert o_fwd.stride(0) == dout.stride(0)
    check(_L.alpro_seq_attn_bwd(_p(qkv), qkv.stride(0), _p(mask), _p(lse), _p(o_fwd), _p(cls_fwd), _p(cls_weight), _p(dout),
                                dout.stride(0), _p(dqkv),
                                _p(scratch), S, nseq, heads, _fmt(qkv), seq_div, stride, clip_rows, scale, drop_p,
                                drop_seed & 0xffffffff, _s()), "alpro_seq_attn_bwd")


def small_linear_fwd(x, ldx, W, b, y, M, N, K, alpha=1.0, alpha_dev=None, alpha_mode=0, relu=False, ldw=None, ldy=None):
    check(_L.alpro_small_linear_fwd(_p(x), ldx, _p(W), ldw if ldw is not None else K, _p(b), _p(y),
                                    ldy if ldy is not None else N, M, N, K, alpha, _p(alpha_dev), alpha_mode,
                                    int(relu), _s()), "alpro_small_linear_fwd")


def small_linear_bwd(dy, lddy, yact, x, ldx, W, dx, lddx, dx_acc, dW, db, dw_acc, M, N, K, alpha=1.0, alpha_dev=None,
                     alpha_mode=0, dw_scale=1.0, ldw=None, lddw=None):
    check(_L.alpro_small_linear_bwd(_p(dy), lddy, _p(yact), lddy, _p(x), ldx, _p(W), ldw if ldw is not None else K,
                                    _p(dx), lddx, int(dx_acc), _p(dW), lddw if lddw is not None else K, _p(db),
                                    int(dw_acc), M, N, K, alpha, _p(alpha_dev), alpha_mode, dw_scale, _s()),
          "alpro_small_linear_bwd")


def l2norm_fwd(x, y, norm, eps=1e-12):
    M, d = x.shape
    check(_L.alpro_l2norm_fwd(_p(x), _p(y), _p(norm), M, d, eps, _s()), "alpro_l2norm_fwd")


def l2norm_bwd(dy, y, norm, dx):
    M, d = y.shape
    check(_L.alpro_l2norm_bwd(_p(dy), _p(y), _p(norm), _p(dx), M, d, _s()), "alpro_l2norm_bwd")


class CEState:
    """Per-row statistics of one softmax-CE evaluation (kept for the backward)."""

    def __init__(self, R, device):
        self.buf = torch.empty(4, R, device=device, dtype=torch.float32)
        self.scal = torch.empty(2, device=device, dtype=torch.float32)  # loss, denom

    row_loss = property(lambda s: s.buf[0])
    row_lse = property(lambda s: s.buf[1])
    row_valid = property(lambda s: s.buf[2])
    row_tsum = property(lambda s: s.buf[3])
    loss = property(lambda s: s.scal[0])
    denom = property(lambda s: s.scal[1:2])


def softmax_ce_fwd(logits, C, hard=None, soft=None, row_ignore=None, denom_mode=0):
    R = logits.shape[0]
    st = CEState(R, logits.device)
    check(_L.alpro_softmax_ce_fwd(_p(logits), logits.stride(0), R, C, _p(hard), _p(soft),
                                  soft.stride(0) if soft is not None else 0, _p(row_ignore), _p(st.buf[0]),
                                  _p(st.buf[1]), _p(st.buf[2]), _p(st.buf[3]), denom_mode, _p(st.scal[0:1]),
                                  _p(st.scal[1:2]), _s()), "alpro_softmax_ce_fwd")
    return st


def softmax_ce_bwd(logits, C, st, gptr, gscale, hard=None, soft=None, out32=None, out16=None, C_out=None):
    R = logits.shape[0]
    out = out32 if out32 is not None else out16
    check(_L.alpro_softmax_ce_bwd(_p(logits), logits.stride(0), R, C, _p(hard), _p(soft),
                                  soft.stride(0) if soft is not None else 0, _p(st.buf[1]), _p(st.buf[2]),
                                  _p(st.buf[3]), _p(st.scal[1:2]), _p(gptr), gscale, _p(out32), _p(out16),
                                  _fmt(out16) if out16 is not None else 0, out.stride(0),
                                  C_out if C_out is not None else C, _s()), "alpro_softmax_ce_bwd")


def temp_grad(dsa, sa, dsb, sb, temp, dtemp, coef):
    check(_L.alpro_temp_grad(_p(dsa), _p(sa), dsa.numel(), _p(dsb), _p(sb), dsb.numel() if dsb is not None else 0,
                             _p(temp), _p(dtemp), coef, _s()), "alpro_temp_grad")


def clamp_scalar(p, lo, hi):
    check(_L.alpro_clamp_scalar(_p(p), lo, hi, _s()), "alpro_clamp_scalar")


def masked_mean_fwd(x, seq_stride, row0, patch_mask, B, Np, h, out):
    check(_L.alpro_masked_mean_fwd(_p(x), seq_stride, row0, _p(patch_mask), B, Np, h, _p(out), _s()),
          "alpro_masked_mean_fwd")


def masked_mean_bwd(dout, patch_mask, B, Np, h, dx, seq_stride, row0):
    check(_L.alpro_masked_mean_bwd(_p(dout), _p(patch_mask), B, Np, h, _p(dx), seq_stride, row0, _s()),
          "alpro_masked_mean_bwd")


def take_rows_fwd(src, R, s0, n, L, h, out32=None, out16=None):
    check(_L.alpro_take_rows_fwd(_p(src), R, s0, n, L, h, _p(out32), _p(out16),
                                 _fmt(out16) if out16 is not None else 0, _s()), "alpro_take_rows_fwd")


def take_rows_bwd(dout, R, s0, n, L, h, dsrc):
    check(_L.alpro_take_rows_bwd(_p(dout), R, s0, n, L, h, _p(dsrc), _s()), "alpro_take_rows_bwd")


def neg_weights(sim, col0, b, w):
    check(_L.alpro_neg_weights(_p(sim), sim.stride(0), col0, b, _p(w), _s()), "alpro_neg_weights")


def neg_sample(sim, col0, b, seed, draw, idx, w=None):
    check(_L.alpro_neg_sample(_p(sim), sim.stride(0), col0, b, seed & 0xffffffff, (seed >> 32) & 0xffffffff,
                              draw & 0xffffffff, _p(w), _p(idx), _s()), "alpro_neg_sample")


def philox4x32_10(ctr_key, out):
    check(_L.alpro_philox4x32_10(_p(ctr_key), _p(out), ctr_key.shape[0], _s()), "alpro_philox4x32_10")


def gelu_grad_mul(dy32, pre16, out16):
    check(_L.alpro_gelu_grad_mul(_p(dy32), _p(pre16), _fmt(pre16), _p(out16), _fmt(out16), dy32.numel(), _s()),
          "alpro_gelu_grad_mul")


def pseudo_labels(sim, soft, ignore):
    R, C = sim.shape
    check(_L.alpro_pseudo_labels(_p(sim), R, C, _p(soft), _p(ignore), _s()), "alpro_pseudo_labels")


def sumsq(x, out):
    check(_L.alpro_sumsq(_p(x), x.numel(), _p(out), _s()), "alpro_sumsq")


def adamw_step(p, g, m, v, beta1, beta2, eps, step_size, lr_wd, gnorm_sq, max_norm):
    check(_L.alpro_adamw_step(_p(p), _p(g), _p(m), _p(v), p.numel(), beta1, beta2, eps, step_size, lr_wd,
                              _p(gnorm_sq), max_norm, _s()), "alpro_adamw_step")


def adamw_prepare(gnorm_sq, lr, beta1, beta2, correct_bias, step_count, step_size_out):
    check(_L.alpro_adamw_prepare(_p(gnorm_sq), lr, beta1, beta2, int(correct_bias), _p(step_count), _p(step_size_out),
                                 _s()), "alpro_adamw_prepare")


def adamw_step_dev(p, g, m, v, beta1, beta2, eps, step_size_dev, lr_wd, gnorm_sq, max_norm):
    check(_L.alpro_adamw_step_dev(_p(p), _p(g), _p(m), _p(v), p.numel(), beta1, beta2, eps, _p(step_size_dev), lr_wd,
                                  _p(gnorm_sq), max_norm, _s()), "alpro_adamw_step_dev")


def nvl_allreduce(peer_ptrs, mc_ptr, world, rank, offset, count, scale, num_ctas=16):
    """peer_ptrs: list of `world` integer device addresses (peer-mapped bases); mc_ptr: multicast address or 0."""
    arr = (ctypes.c_void_p * world)(*[ctypes.c_void_p(int(p)) for p in peer_ptrs])
    check(_L.alpro_nvl_allreduce(ctypes.cast(arr, ctypes.c_void_p), int(mc_ptr) if mc_ptr else None, world, rank,
                                 offset, count, scale, num_ctas, _s()), "alpro_nvl_allreduce")


def sum_slices(own, stage, nparts, count, part_stride, scale):
    check(_L.alpro_sum_slices(_p(own), _p(stage), nparts, count, part_stride, scale, _s()), "alpro_sum_slices")


def memcpy_async(dst_ptr, src_ptr, nbytes):
    check(_L.alpro_memcpy_async(int(dst_ptr), int(src_ptr), nbytes, _s()), "alpro_memcpy_async")
