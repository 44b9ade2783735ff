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


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


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
           out16=None, out16b=None, skip_period=0, split_k=0, alpha=1.0):
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
    check(_lib.alpro_gemm16(_ptr(a), _ptr(b), M, N, K, lda, ldb, a_layout, b_layout, _fmt(a), _fmt(b),
                            ctypes.byref(ep), _stream()), "alpro_gemm16")
