"""Collectives of the hot path over torch.distributed (NCCL on NVLink 5 / NVSwitch on the GPU box, gloo in CPU tests).

Replaces the Horovod calls of the reference:
  * hvd.allgather of the normalised VTC features (alpro_models.py:110-111, 565-566, 764-765): forward = concatenation in
    rank order; backward = sum over ranks of the gathered gradient, narrowed to the local slice. We issue it as a
    reduce-scatter (half the bytes of Horovod's allreduce + narrow, identical values).
  * hvd.DistributedOptimizer gradient averaging (run_video_retrieval.py:320-323, 444): ONE all-reduce (op=AVG) over the
    flat gradient buffer the backward already writes into (GradStore.flat) — a single NVLS-sized message instead of
    ~300 per-tensor reductions; frozen teacher / unused parameters are not part of the buffer.
"""
import torch
import torch.distributed as dist


class TorchDistComm:
    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._gloo = dist.get_backend(group) == "gloo"

    def all_gather(self, x):
        x = x.contiguous()
        out = torch.empty((self.world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x, group=self.group)
        return out

    def reduce_scatter_sum(self, g):
        g = g.contiguous()
        b = g.shape[0] // self.world
        if self._gloo:  # gloo has no reduce_scatter: all-reduce + narrow (what Horovod's allgather grad does)
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
            return g[self.rank * b:(self.rank + 1) * b].clone()
        out = torch.empty((b,) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
        dist.reduce_scatter_tensor(out, g, op=dist.ReduceOp.SUM, group=self.group)
        return out


class BucketedAllReduce:
    """Gradient averaging overlapped with the backward pass.

    The backward writes all gradients into one flat buffer whose layout follows completion order (GradStore.regions);
    the engine calls `ready(G, end)` whenever the prefix [0, end) is final. Ready slices are all-reduced (AVG) on a side
    stream in buckets of at least `min_bucket` elements while the remaining TimeSformer blocks are still running their
    backward; `finish()` (called by allreduce_gradients) flushes the tail and joins the streams. Replaces the blocking
    `optimizer.synchronize()` of hvd.DistributedOptimizer (run_video_retrieval.py:444)."""

    def __init__(self, group=None, min_bucket=32 * 1024 * 1024):
        self.group = group
        self.min_bucket = min_bucket
        self.world = dist.get_world_size(group)
        self._gloo = dist.get_backend(group) == "gloo"
        self.side = torch.cuda.Stream() if torch.cuda.is_available() and not self._gloo else None
        self._done = 0
        self._G = None
        self.bytes = 0

    def _reduce(self, flat, lo, hi):
        if hi <= lo:
            return
        sl = flat[lo:hi]
        self.bytes += (hi - lo) * 4
        if self._gloo:
            dist.all_reduce(sl, op=dist.ReduceOp.SUM, group=self.group)
            sl.div_(self.world)
            return
        ev = torch.cuda.Event()
        ev.record()                                   # gradients of [lo, hi) are complete on the compute stream here
        with torch.cuda.stream(self.side):
            self.side.wait_event(ev)
            dist.all_reduce(sl, op=dist.ReduceOp.AVG, group=self.group)

    def ready(self, G, end):
        if self._G is not G:                          # new backward pass
            self._G, self._done, self.bytes = G, 0, 0
        final = end >= G.flat.numel()
        if end - self._done >= self.min_bucket or final:
            self._reduce(G.flat, self._done, end)
            self._done = end

    def finish(self):
        G = self._G
        if G is None:
            raise RuntimeError("no gradients to reduce: run loss.backward() first")
        if self._done < G.flat.numel():
            self._reduce(G.flat, self._done, G.flat.numel())
            self._done = G.flat.numel()
        if self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)
        self._G = None
        return self.bytes

    def abandon(self):
        """Join the side stream and forget the pending store (its p.grad tensors were reduced another way)."""
        if self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)
        self._G = None


def attach(model, group=None, overlap=True):
    """Make `model` exchange VTC features across the ranks of `group` and (overlap=True) average its gradients while
    the backward pass is still running (call after dist.init_process_group)."""
    model.engine.comm = TorchDistComm(group)
    if overlap:
        model._grad_reducer = BucketedAllReduce(group)
        model.engine.grad_ready_hook = model._grad_reducer.ready
    return model


def allreduce_gradients(model, group=None):
    """Average all parameter gradients across ranks. With comm.attach(model, overlap=True) most of the buffer has already
    been reduced on the side stream during backward and only the tail + stream join happen here; otherwise one collective
    over the flat gradient buffer."""
    world = dist.get_world_size(group)
    if not getattr(model, "_grads_aliased", True):
        # p.grad was assigned by hand and does not alias the flat store: reduce the p.grad tensors themselves
        gs = [p.grad for p in model.parameters() if p.grad is not None]
        if not gs:
            raise RuntimeError("no gradients to reduce: run loss.backward() first")
        flat = torch.cat([g.reshape(-1) for g in gs])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        o = 0
        for g in gs:
            g.copy_(flat[o:o + g.numel()].view_as(g))
            o += g.numel()
        red = getattr(model, "_grad_reducer", None)
        if red is not None and red._G is not None:
            red.abandon()
        return flat.numel() * 4
    red = getattr(model, "_grad_reducer", None)
    if red is not None:
        # overlap=True: everything but the tail was reduced during backward; nothing pending = already reduced
        # (gradient accumulation joins each micro-step's reduction inside backward, modeling._publish_grads)
        return red.finish() if red._G is not None else 0
    flat = getattr(model.engine, "last_grads", None)
    if flat is None:
        raise RuntimeError("no gradients to reduce: run loss.backward() first")
    if dist.get_backend(group) == "gloo":
        dist.all_reduce(flat.flat, op=dist.ReduceOp.SUM, group=group)
        flat.flat.div_(world)
    else:
        dist.all_reduce(flat.flat, op=dist.ReduceOp.AVG, group=group)
    return flat.flat.numel() * 4
