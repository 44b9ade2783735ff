"""ctypes binding of libalpro_b200.so (the C-ABI declared in include/alpro_b200.h).

The product path has no fallback: if the shared library is missing or a symbol cannot be resolved, import fails loudly.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libalpro_b200.so")

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_int64 = ctypes.c_int64
c_float = ctypes.c_float


class GemmEpilogue(ctypes.Structure):
    """Mirror of AlproGemmEpilogue (include/alpro_b200.h)."""
    _fields_ = [
        ("bias", c_void_p), ("aux16", c_void_p), ("resid", c_void_p),
        ("out32", c_void_p), ("out16", c_void_p), ("out16b", c_void_p),
        ("ld32", c_int64), ("ld16", c_int64), ("ld16b", c_int64), ("ldresid", c_int64), ("ldaux", c_int64),
        ("out16_fmt", ctypes.c_int32), ("out16b_fmt", ctypes.c_int32), ("aux_fmt", ctypes.c_int32),
        ("act", ctypes.c_int32), ("skip_period", ctypes.c_int32), ("split_k", ctypes.c_int32),
        ("alpha", c_float),
    ]


class AlproError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m alpro_b200.build` (or __graft_entry__.build()). "
            "alpro_b200 has no CPU / PyTorch fallback for its kernels.")
    return ctypes.CDLL(LIB_PATH)


lib = _load()

# name -> (restype, argtypes); every symbol declared in include/alpro_b200.h must be listed here
# (tests/test_abi.py cross-checks this table against the header).
_SIGS = {}


def _sig(name, argtypes, restype=c_int):
    fn = getattr(lib, name)  # AttributeError if the symbol is missing -> loud failure
    fn.restype = restype
    fn.argtypes = argtypes
    _SIGS[name] = fn
    return fn


alpro_last_error = _sig("alpro_last_error", [], ctypes.c_char_p)
alpro_version = _sig("alpro_version", [])
alpro_num_sms = _sig("alpro_num_sms", [])
alpro_gemm16 = _sig("alpro_gemm16", [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64,
                                     c_int, c_int, c_int, c_int, ctypes.POINTER(GemmEpilogue), c_void_p])


def check(rc, what=""):
    if rc != 0:
        msg = alpro_last_error()
        raise AlproError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")
