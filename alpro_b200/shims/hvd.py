"""`horovod.torch` stand-in over torch.distributed (NCCL on the GPU box, gloo on CPU), covering exactly the surface the
reference's run scripts and utilities use:

    init / rank / size / local_rank / local_size                      run_video_retrieval.py:305-313, 818
    allgather(t, name=None)      concat along dim 0 in rank order     src/utils/distributed.py:168,234-235
    allreduce_(t) / allreduce(t) in-place / out-of-place AVERAGE      src/utils/distributed.py:35,68,83
    broadcast_(t, root_rank) / broadcast_parameters / broadcast_optimizer_state
                                                                      run_video_retrieval.py:326-327
    DistributedOptimizer(opt, named_parameters=, compression=) with synchronize() / skip_synchronize()
                                                                      run_video_retrieval.py:320-323,444,486-488
    Compression.none / Compression.fp16

Gradient averaging of an alpro_b200 model goes through alpro_b200.comm.allreduce_gradients (one flat buffer, reduced in
buckets under the backward pass); parameters that do not belong to such a model are averaged tensor by tensor.
Installed as `horovod.torch` by alpro_b200.shims.install() only when the real Horovod is not importable.
"""
import contextlib
import os

import torch
import torch.distributed as dist

_state = {"init": False}


def init(comm=None):
    """hvd.init(): under torchrun (RANK / WORLD_SIZE in the environment) joins the process group, NCCL when CUDA is
    visible; otherwise single process."""
    if not dist.is_initialized() and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if torch.cuda.is_available():
            torch.cuda.set_device(local_rank())
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank()))
        else:
            dist.init_process_group("gloo")
    _state["init"] = True


def shutdown():
    if dist.is_initialized():
        dist.destroy_process_group()
    _state["init"] = False


def is_initialized():
    return _state["init"]


def _multi():
    return dist.is_initialized() and dist.get_world_size() > 1


def rank():
    return dist.get_rank() if dist.is_initialized() else int(os.environ.get("RANK", "0"))


def size():
    return dist.get_world_size() if dist.is_initialized() else 1


def local_rank():
    return int(os.environ.get("LOCAL_RANK", rank()))


def local_size():
    return int(os.environ.get("LOCAL_WORLD_SIZE", str(size())))


def _dev(t):
    """NCCL needs CUDA tensors; gloo works on the tensor's own device."""
    if dist.is_initialized() and dist.get_backend() == "nccl" and not t.is_cuda:
        return t.cuda()
    return t


class _AllGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _dev(x.contiguous())
        world = dist.get_world_size()
        n = torch.tensor([x.shape[0]], dtype=torch.int64, device=x.device)
        ns = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(ns, n)
        ns = [int(v) for v in ns]
        ctx.ns, ctx.rank = ns, dist.get_rank()
        mx = max(ns)
        pad = x if x.shape[0] == mx else torch.cat([x, x.new_zeros((mx - x.shape[0],) + tuple(x.shape[1:]))])
        out = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(out, pad)
        return torch.cat([o[:k] for o, k in zip(out, ns)], dim=0)

    @staticmethod
    def backward(ctx, g):
        # Horovod 0.19.4 allgather gradient: sum over ranks, then the local slice
        g = g.contiguous().clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        lo = sum(ctx.ns[:ctx.rank])
        return g[lo:lo + ctx.ns[ctx.rank]]


def allgather(tensor, name=None):
    if not _multi():
        return tensor
    out = _AllGather.apply(tensor)
    return out if tensor.is_cuda or not out.is_cuda else out.cpu()


def allreduce_(tensor, average=True, name=None, op=None):
    if _multi():
        t = _dev(tensor)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        if average:
            t.div_(dist.get_world_size())
        if t is not tensor:
            tensor.copy_(t)
    return tensor


def allreduce(tensor, average=True, name=None, op=None):
    return allreduce_(tensor.clone(), average=average, name=name)


def broadcast_(tensor, root_rank, name=None):
    if _multi():
        t = _dev(tensor)
        dist.broadcast(t, src=root_rank)
        if t is not tensor:
            tensor.copy_(t)
    return tensor


def broadcast(tensor, root_rank, name=None):
    return broadcast_(tensor.clone(), root_rank, name)


def broadcast_parameters(params, root_rank=0):
    """params: a state_dict or an iterable of (name, tensor)."""
    if not _multi():
        return
    items = params.items() if isinstance(params, dict) else params
    for _, p in sorted(items, key=lambda kv: kv[0]):
        if torch.is_tensor(p):
            broadcast_(p.data if isinstance(p, torch.nn.Parameter) else p, root_rank)


def broadcast_optimizer_state(optimizer, root_rank=0):
    if not _multi():
        return
    obj = [optimizer.state_dict() if dist.get_rank() == root_rank else None]
    dist.broadcast_object_list(obj, src=root_rank)
    if dist.get_rank() != root_rank:
        optimizer.load_state_dict(obj[0])


def broadcast_object(obj, root_rank=0, name=None):
    if not _multi():
        return obj
    box = [obj]
    dist.broadcast_object_list(box, src=root_rank)
    return box[0]


class Compression:
    class none:
        @staticmethod
        def compress(t):
            return t, None

        @staticmethod
        def decompress(t, ctx):
            return t

    class fp16:
        @staticmethod
        def compress(t):
            return (t.half(), t.dtype) if t.is_floating_point() else (t, None)

        @staticmethod
        def decompress(t, ctx):
            return t.to(ctx) if ctx is not None else t


# ---------------------------------------------------------------------------------------------------- optimizer
def _owning_models(params):
    """alpro_b200 models (registered at construction) that own any of `params`."""
    from .. import modeling
    ids = {id(p) for p in params}
    owners = []
    for m in modeling.live_models():
        mine = {id(p) for p in m.parameters()}
        if mine & ids:
            owners.append((m, mine))
    return owners


class _DistributedOptimizerMixin:
    def _hvd_setup(self, named_parameters, compression, backward_passes_per_step):
        self._hvd_compression = compression
        self._hvd_synchronized = False
        self._hvd_should_sync = True
        self._hvd_named = list(named_parameters) if named_parameters is not None else []

    def synchronize(self):
        """Average the gradients over ranks (hvd.DistributedOptimizer.synchronize, run_video_retrieval.py:444)."""
        if _multi():
            from .. import comm
            params = [p for g in self.param_groups for p in g["params"]]
            covered = set()
            for model, mine in _owning_models(params):
                comm.allreduce_gradients(model)
                covered |= mine
            world = dist.get_world_size()
            for p in params:
                if id(p) not in covered and p.grad is not None:
                    c, ctx = self._hvd_compression.compress(p.grad)
                    dist.all_reduce(c, op=dist.ReduceOp.SUM)
                    p.grad.copy_(self._hvd_compression.decompress(c, ctx)).div_(world)
        self._hvd_synchronized = True

    @contextlib.contextmanager
    def skip_synchronize(self):
        self._hvd_should_sync = False
        try:
            yield
        finally:
            self._hvd_should_sync = True

    def step(self, closure=None):
        if self._hvd_should_sync and not self._hvd_synchronized:
            self.synchronize()
        self._hvd_synchronized = False
        return super(self.__class__, self).step(closure) if closure is not None else super(self.__class__, self).step()


def DistributedOptimizer(optimizer, named_parameters=None, compression=Compression.none, backward_passes_per_step=1,
                         op=None):
    """Same construction as Horovod's: a dynamic subclass of the wrapped optimizer's class sharing its param_groups."""
    cls = type(optimizer.__class__.__name__, (optimizer.__class__,),
               {k: v for k, v in _DistributedOptimizerMixin.__dict__.items() if not k.startswith("__")})
    new = cls.__new__(cls)
    new.__dict__.update(optimizer.__dict__)
    new._hvd_setup(named_parameters, compression, backward_passes_per_step)
    return new
