"""Import-time dependencies of the UNCHANGED reference run scripts (src/tasks/run_video_retrieval.py:9-37,
src/pretrain/run_pretrain_sparse.py:1-33) that a B200 box does not need or have:

  * `horovod.torch`  -> alpro_b200.shims.hvd  (torch.distributed / NCCL over NVLink)
  * `apex.amp`       -> alpro_b200.shims.amp  (pass-through; the kernels own their operand format)
  * `apex.normalization.fused_layer_norm.FusedLayerNorm` -> torch.nn.LayerNorm (imported by src/modeling/xbert.py:46 only)
  * `src.modeling.alpro_models` -> alpro_b200.modeling, so `from src.modeling.alpro_models import AlproForPretrain`
    in the run scripts resolves to the B200 classes (install(alias_models=True))

install() registers a stand-in ONLY when the real package cannot be imported. `python -m alpro_b200.launch
src/tasks/run_video_retrieval.py --config ...` calls it and then runs the script unmodified.
"""
import importlib
import importlib.util
import sys
import types


def _missing(name):
    if name in sys.modules:
        return False
    try:
        return importlib.util.find_spec(name) is None
    except (ImportError, ValueError, AttributeError):
        return True


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__alpro_shim__ = True
    sys.modules[name] = m
    return m


def install(alias_models=True, optional_stubs=False, force=False):
    """Returns the list of module names that were provided by a stand-in. force=True replaces whatever `horovod` /
    `apex` modules are registered (tests use it to displace other test doubles)."""
    provided = []
    if force:
        for k in [k for k in sys.modules if k.split(".")[0] in ("horovod", "apex")]:
            del sys.modules[k]
    if force or _missing("horovod"):
        from . import hvd
        pkg = _module("horovod", __path__=[])
        sys.modules["horovod.torch"] = hvd
        pkg.torch = hvd
        if not hasattr(hvd, "__path__"):
            hvd.__path__ = []           # `from horovod.torch.mpi_ops import rank, size` (src/utils/distributed.py:14)
        ops_mod = _module("horovod.torch.mpi_ops", **{k: getattr(hvd, k) for k in (
            "init", "shutdown", "rank", "size", "local_rank", "local_size", "allgather", "allreduce", "allreduce_",
            "broadcast", "broadcast_")})
        hvd.mpi_ops = ops_mod
        provided += ["horovod", "horovod.torch", "horovod.torch.mpi_ops"]
    if force or _missing("apex"):
        import torch
        from . import amp
        pkg = _module("apex", __path__=[])
        sys.modules["apex.amp"] = amp
        pkg.amp = amp
        norm = _module("apex.normalization", __path__=[])
        fln = _module("apex.normalization.fused_layer_norm", FusedLayerNorm=torch.nn.LayerNorm)
        norm.fused_layer_norm = fln
        pkg.normalization = norm
        provided += ["apex", "apex.amp", "apex.normalization.fused_layer_norm"]
    if optional_stubs:
        provided += _optional_stubs()
    if alias_models:
        from .. import modeling
        sys.modules["src.modeling.alpro_models"] = modeling
        provided.append("src.modeling.alpro_models")
    return provided


def _optional_stubs():
    """Small pure-Python conveniences of the run scripts, provided only when absent: easydict.EasyDict, ujson (= json),
    tensorboardX.SummaryWriter (torch's writer or a no-op). Data-decoding libraries (decord, av, lmdb) are NOT faked:
    a box that trains from video files has them installed."""
    import json
    got = []
    if _missing("easydict"):
        class EasyDict(dict):
            def __init__(self, d=None, **kw):
                super().__init__()
                for k, v in dict(d or {}, **kw).items():
                    self[k] = v

            def __setitem__(self, k, v):
                if isinstance(v, dict) and not isinstance(v, EasyDict):
                    v = EasyDict(v)
                elif isinstance(v, (list, tuple)):
                    v = type(v)(EasyDict(x) if isinstance(x, dict) else x for x in v)
                super().__setitem__(k, v)

            __setattr__ = __setitem__

            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError:
                    raise AttributeError(k)
        _module("easydict", EasyDict=EasyDict)
        got.append("easydict")
    if _missing("ujson"):
        _module("ujson", **{k: getattr(json, k) for k in ("load", "loads", "dump", "dumps")})
        got.append("ujson")
    if _missing("tensorboardX"):
        try:
            from torch.utils.tensorboard import SummaryWriter
        except Exception:
            class SummaryWriter:
                def __init__(self, *a, **k):
                    pass

                def __getattr__(self, name):
                    return lambda *a, **k: None
        _module("tensorboardX", SummaryWriter=SummaryWriter)
        got.append("tensorboardX")
    return got
