"""ctypes binding of libalpro_b200.so. Signatures are read from include/alpro_b200.h (the single source of truth for
the C-ABI), so every declared entry point must be exported by the library or the import fails loudly.

The product path has no fallback: if the shared library is missing or a symbol cannot be resolved, import fails.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libalpro_b200.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "alpro_b200.h")

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
        ("row_scale_acc", c_void_p), ("row_scale_bias", c_void_p),
        ("bias2", c_void_p),
    ]


class AlproError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m alpro_b200.build` (or __graft_entry__.build()). "
            "alpro_b200 has no CPU / PyTorch fallback for its kernels.")
    return ctypes.CDLL(LIB_PATH)


def parse_header(path=HEADER_PATH):
    """Returns {name: (restype, [argtypes])} for every function prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    protos = {}
    for m in re.finditer(r"\b(int|const char\s*\*)\s+(alpro_\w+)\s*\(([^;{]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        restype = c_int if ret == "int" else ctypes.c_char_p
        argtypes = []
        args = args.strip()
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "AlproGemmEpilogue" in a:
                    argtypes.append(ctypes.POINTER(GemmEpilogue))
                elif "*" in a:
                    argtypes.append(c_void_p)
                elif re.match(r"(const )?int64_t\b", a):
                    argtypes.append(c_int64)
                elif re.match(r"(const )?uint32_t\b", a):
                    argtypes.append(ctypes.c_uint32)
                elif re.match(r"(const )?(int|int32_t)\b", a):
                    argtypes.append(c_int)
                elif re.match(r"(const )?float\b", a):
                    argtypes.append(c_float)
                else:
                    raise ImportError(f"alpro_b200.h: cannot map argument '{a}' of {name}")
        protos[name] = (restype, argtypes)
    return protos


lib = _load()
PROTOS = parse_header()
for _name, (_res, _args) in PROTOS.items():
    _fn = getattr(lib, _name)  # AttributeError if the symbol is missing -> loud failure
    _fn.restype = _res
    _fn.argtypes = _args
    globals()[_name] = _fn

_NO_LAUNCH = {"alpro_last_error", "alpro_version", "alpro_num_sms", "alpro_comm_unique_id", "alpro_comm_init",
              "alpro_comm_destroy", "alpro_comm_rank", "alpro_comm_world"}


class _Counting:
    """Attribute proxy over the CDLL that counts kernel-launching C-ABI calls (bench.py reports it as gpu_launches)."""

    def __init__(self, cdll):
        self._cdll = cdll
        self.calls = 0
        for name in PROTOS:
            fn = getattr(cdll, name)
            if name in _NO_LAUNCH:
                setattr(self, name, fn)
            else:
                setattr(self, name, self._wrap(fn))

    def _wrap(self, fn):
        def call(*a):
            self.calls += 1
            return fn(*a)
        return call


counted = _Counting(lib)


def check(rc, what=""):
    if rc != 0:
        msg = lib.alpro_last_error()
        raise AlproError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")
