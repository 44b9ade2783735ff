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


class NcclAbiComm:
    """The same exchange through the library's own C-ABI communicator (alpro_comm_*, include/alpro_b200.h) instead of
    torch.distributed — what a non-Python host binds. `id128` comes from NcclAbiComm.unique_id() on rank 0, shipped to
    the other ranks out of band (file, socket, torch.distributed.broadcast_object_list ...)."""

    def __init__(self, world, rank, id128):
        import ctypes
        from . import _lib
        self._lib, self._ct = _lib, ctypes
        self.world, self.rank = world, rank
        h = ctypes.c_void_p()
        buf = (ctypes.c_char * 128).from_buffer_copy(bytes(id128))
        _lib.check(_lib.lib.alpro_comm_init(ctypes.byref(h), world, rank, ctypes.cast(buf, ctypes.c_void_p)),
                   "alpro_comm_init")
        self._h = h

    @staticmethod
    def unique_id():
        import ctypes
        from . import _lib
        buf = (ctypes.c_char * 128)()
        _lib.check(_lib.lib.alpro_comm_unique_id(ctypes.cast(buf, ctypes.c_void_p)), "alpro_comm_unique_id")
        return bytes(buf)

    _KIND = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}

    def _stream(self):
        return torch.cuda.current_stream().cuda_stream

    def all_gather(self, x):
        x = x.contiguous()
        out = torch.empty((self.world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        self._lib.check(self._lib.counted.alpro_comm_allgather(self._h, x.data_ptr(), out.data_ptr(), x.numel(),
                                                               self._KIND[x.dtype], self._stream()), "alpro_comm_allgather")
        return out

    def reduce_scatter_sum(self, g):
        g = g.contiguous()
        b = g.shape[0] // self.world
        out = torch.empty((b,) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
        self._lib.check(self._lib.counted.alpro_comm_reduce_scatter(self._h, g.data_ptr(), out.data_ptr(), out.numel(),
                                                                    self._KIND[g.dtype], self._stream()),
                        "alpro_comm_reduce_scatter")
        return out

    def all_reduce_(self, t, average=True):
        assert t.is_contiguous()
        self._lib.check(self._lib.counted.alpro_comm_allreduce(self._h, t.data_ptr(), t.numel(), self._KIND[t.dtype],
                                                               int(average), self._stream()), "alpro_comm_allreduce")
        return t

    def close(self):
        if self._h:
            self._lib.lib.alpro_comm_destroy(self._h)
            self._h = None


class BucketedAllReduce:
    """Gradient averaging overlapped with the backward pass.

    The backward writes all gradients into one flat buffer whose layout follows completion order (GradStore.regions);
    the engine calls `ready(G, end)` whenever the prefix [0, end) is final. Ready slices are all-reduced (AVG) on a side
    stream in buckets of at least `min_bucket` elements while the remaining TimeSformer blocks are still running their
    backward; `finish()` (called by allreduce_gradients) flushes the tail and joins the streams. Replaces the blocking
    `optimizer.synchronize()` of hvd.DistributedOptimizer (run_video_retrieval.py:444)."""

    def __init__(self, group=None, min_bucket=32 * 1024 * 1024, compress=None, overlap=True):
        """compress='bf16': each bucket crosses NVLink as bf16 (scaled by 1/world before rounding, summed in bf16,
        widened back into the fp32 store) — half the bytes of the reference's fp32 averaging, gradient error ~2^-9
        relative; off by default (hvd.Compression.none, run_video_retrieval.py:320). overlap=False queues the buckets
        on the compute stream instead of a side stream (no kernel concurrency with the backward GEMMs)."""
        self.group = group
        self.min_bucket = min_bucket
        self.world = dist.get_world_size(group)
        self._gloo = dist.get_backend(group) == "gloo"
        self.compress = compress
        self.overlap = overlap
        self.side = torch.cuda.Stream() if torch.cuda.is_available() and not self._gloo and overlap else None
        self._done = 0
        self._G = None
        self.bytes = 0
        self._stage = None

    def _reduce_on_stream(self, sl):
        if self.compress == "bf16":
            if self._stage is None or self._stage.numel() < sl.numel():
                self._stage = torch.empty(max(sl.numel(), self.min_bucket), dtype=torch.bfloat16, device=sl.device)
            st = self._stage[:sl.numel()]
            torch.mul(sl, 1.0 / self.world, out=st)
            dist.all_reduce(st, op=dist.ReduceOp.SUM, group=self.group)
            sl.copy_(st)
        else:
            dist.all_reduce(sl, op=dist.ReduceOp.AVG, group=self.group)

    def _reduce(self, flat, lo, hi):
        if hi <= lo:
            return
        sl = flat[lo:hi]
        self.bytes += (hi - lo) * (2 if self.compress == "bf16" else 4)
        if self._gloo:
            dist.all_reduce(sl, op=dist.ReduceOp.SUM, group=self.group)
            sl.div_(self.world)
            return
        if self.side is None:
            self._reduce_on_stream(sl)
            return
        ev = torch.cuda.Event()
        ev.record()                                   # gradients of [lo, hi) are complete on the compute stream here
        with torch.cuda.stream(self.side):
            self.side.wait_event(ev)
            self._reduce_on_stream(sl)

    def ready(self, G, end):
        if self._G is not G:                          # new backward pass
            self._G, self._done, self.bytes = G, 0, 0
        n = G.flat.numel()
        final = end >= n
        # towards the end of the backward the buckets shrink: what is still pending when the last gradient lands is the
        # EXPOSED part of the reduction (pull + sum + push + barrier of the tail bucket)
        thr = self.min_bucket if end < 0.75 * n else self.min_bucket // 4
        if end - self._done >= thr or final:
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


class NvlGradReducer(BucketedAllReduce):
    """Gradient averaging by OUR kernel over NVLink peer memory (csrc/allreduce.cu) instead of ncclAllReduce.

    The engine's flat gradient store lives in symmetric memory (torch.distributed._symmetric_memory: cuMem allocation
    mapped into every rank of the node, plus an NVSwitch multicast mapping when the fabric offers one). A ready bucket is
    reduced on the side stream by   barrier -> alpro_nvl_allreduce -> barrier   where the barriers are the symmetric
    memory's signal-pad barriers (one tiny kernel each) and the reduction is a thin kernel — 128-thread CTAs, no shared
    memory, `num_ctas` of them — that shares SMs with the backward GEMMs instead of displacing their CTAs as the NCCL
    kernel does (profiles/r02c_scaling_probe_n2.md). With multicast the adds happen inside the switch
    (multimem.ld_reduce) and each rank moves 1/W of the bucket; otherwise W peer loads / stores per element."""

    def __init__(self, group=None, min_bucket=32 * 1024 * 1024, num_ctas=16):
        super().__init__(group, min_bucket=min_bucket, compress=None, overlap=True)
        import torch.distributed._symmetric_memory as symm_mem
        self._sm = symm_mem
        self.group_obj = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(group)
        self.num_ctas = num_ctas
        self._bufs = []          # [(tensor, handle, peer_ptrs, mc_ptr)]
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            try:                                  # needed by older torch, a deprecated no-op from 2.9 on
                symm_mem.enable_symm_mem_for_group(self.group_obj.group_name)
            except Exception:
                pass

    def alloc(self, numel, device, avoid=None):
        """A zeroed flat fp32 buffer of `numel` elements in symmetric memory. COLLECTIVE on first use (every rank reaches
        its first backward together). `avoid`: a buffer still aliased by live p.grad (gradient accumulation)."""
        for ent in self._bufs:
            t = ent[0]
            if t.numel() == numel and (avoid is None or t.data_ptr() != avoid.data_ptr()):
                t.zero_()
                return t
        t = self._sm.empty(numel, dtype=torch.float32, device=device)
        hdl = self._sm.rendezvous(t, self.group_obj.group_name)
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        mc = int(hdl.multicast_ptr) if getattr(hdl, "has_multicast_support", lambda *a: False) and hdl.multicast_ptr else 0
        self._bufs.append((t, hdl, ptrs, mc))
        t.zero_()
        return t

    def _entry(self, flat):
        for ent in self._bufs:
            if ent[0].data_ptr() == flat.data_ptr():
                return ent
        return None

    def uses_multicast(self):
        return bool(self._bufs) and all(e[3] != 0 for e in self._bufs)

    def _reduce_on_stream(self, sl):
        ent = self._entry(self._G.flat)
        if ent is None:                       # store not in symmetric memory (hand-built GradStore): NCCL
            return super()._reduce_on_stream(sl)
        t, hdl, ptrs, mc = ent
        from . import ops
        off = (sl.data_ptr() - ptrs[self.rank]) // 4      # floats from the base of the symmetric allocation
        hdl.barrier(channel=0)                # every rank's gradients of this bucket are final
        ops.nvl_allreduce(ptrs, mc, self.world, self.rank, off, sl.numel(), 1.0 / self.world, self.num_ctas)
        hdl.barrier(channel=1)                # every rank's slice has been written back everywhere


class CeGradReducer(NvlGradReducer):
    """Gradient averaging with NO resident kernel at all: the bytes move by the copy engines.

    Rank r owns the r-th 1/W of every bucket. Per bucket, on the side stream: barrier (bucket final everywhere), W-1
    cudaMemcpyAsync PULLS of the peers' copies of the owned slice into a staging buffer (NVLink, copy engines). The sum
    `own = (own + sum staged) / W` is an ordinary full-width kernel placed IN ORDER on the compute stream one bucket
    later (its inputs have long arrived, so the compute stream does not stall and nothing runs concurrently with the
    GEMMs); after it, W-1 PUSHES of the averaged slice to the peers, again on the side stream. One closing barrier in
    finish(). Motivation and numbers: profiles/r02c_scaling_probe_n2.md, r02d."""

    def __init__(self, group=None, min_bucket=16 * 1024 * 1024):
        super().__init__(group, min_bucket=min_bucket)
        self._stage = [None, None]
        self._pending = None            # (lo, hi, k, pulled_event) of the bucket whose sum has not been enqueued
        self._summed = {}               # k -> event: sum of bucket k done (staging buffer k % 2 reusable)
        self._k = 0

    def _slice(self, lo, hi):
        n4 = (hi - lo + 3) // 4
        c4 = (n4 + self.world - 1) // self.world
        a = min(n4, self.rank * c4)
        b = min(n4, a + c4)
        return lo + 4 * a, 4 * (b - a), 4 * c4

    def _reduce(self, flat, lo, hi):
        if hi <= lo:
            return
        ent = self._entry(flat)
        if ent is None or self._gloo:
            return BucketedAllReduce._reduce(self, flat, lo, hi)
        from . import ops
        t, hdl, ptrs, mc = ent
        self.bytes += (hi - lo) * 4
        k = self._k
        self._k += 1
        start, n_own, cap = self._slice(lo, hi)
        slot = k & 1
        need = cap * max(self.world - 1, 1)
        if self._stage[slot] is None or self._stage[slot].numel() < need:
            self._stage[slot] = torch.empty(need, dtype=torch.float32, device=flat.device)
        ev = torch.cuda.Event()
        ev.record()                                   # bucket k complete on the compute stream
        # the previous bucket's sum goes onto the compute stream NOW (its pulls were issued one bucket ago)
        self._flush_pending(ptrs, t)
        with torch.cuda.stream(self.side):
            self.side.wait_event(ev)
            if (k - 2) in self._summed:               # staging slot reuse: its last consumer has run
                self.side.wait_event(self._summed.pop(k - 2))
            hdl.barrier(channel=0)                    # every rank's gradients of bucket k are final
            pulled = torch.cuda.Event()
            if n_own > 0:
                off_b = (t.data_ptr() - ptrs[self.rank]) + start * 4     # bytes from the base of the allocation
                j = 0
                for p in range(self.world):
                    if p == self.rank:
                        continue
                    ops.memcpy_async(self._stage[slot].data_ptr() + j * cap * 4, ptrs[p] + off_b, n_own * 4)
                    j += 1
            pulled.record()
        self._pending = (start, n_own, cap, k, pulled, slot)

    def _flush_pending(self, ptrs, t):
        if self._pending is None:
            return
        from . import ops
        start, n_own, cap, k, pulled, slot = self._pending
        self._pending = None
        cur = torch.cuda.current_stream()
        cur.wait_event(pulled)
        if n_own > 0:
            ops.sum_slices(t[start:start + n_own], self._stage[slot], self.world - 1, n_own, cap, 1.0 / self.world)
        done = torch.cuda.Event()
        done.record()
        self._summed[k] = done
        with torch.cuda.stream(self.side):
            self.side.wait_event(done)
            if n_own > 0:
                off_b = (t.data_ptr() - ptrs[self.rank]) + start * 4
                for p in range(self.world):
                    if p != self.rank:
                        ops.memcpy_async(ptrs[p] + off_b, ptrs[self.rank] + off_b, n_own * 4)

    def finish(self):
        G = self._G
        if G is None:
            raise RuntimeError("no gradients to reduce: run loss.backward() first")
        if self._done < G.flat.numel():
            self._reduce(G.flat, self._done, G.flat.numel())
            self._done = G.flat.numel()
        ent = self._entry(G.flat)
        if ent is not None and not self._gloo:
            t, hdl, ptrs, mc = ent
            self._flush_pending(ptrs, t)
            with torch.cuda.stream(self.side):
                hdl.barrier(channel=1)                # every rank's pushes have landed everywhere
            self._summed.clear()
            self._k = 0
        if self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)
        self._G = None
        return self.bytes


def attach(model, group=None, overlap=True):
    """Make `model` exchange VTC features across the ranks of `group` and (overlap=True) average its gradients while
    the backward pass is still running (call after dist.init_process_group)."""
    import os
    model.engine.comm = TorchDistComm(group)
    mode = os.environ.get("ALPRO_DP_OVERLAP", "1" if overlap else "0")    # 1: side stream, 0: one reduce at the end,
    compress = os.environ.get("ALPRO_GRAD_COMPRESS", "none")             # stream: bucketed on the compute stream
    compress = None if compress in ("none", "", "fp32") else compress
    model.engine.grad_alloc = None
    if mode != "0":
        red = None
        # nccl: ncclAllReduce buckets on a side stream (round 1); nvl: our multimem / P2P kernel; ce: copy engines
        backend = os.environ.get("ALPRO_GRAD_REDUCER", "ce")
        if backend in ("nvl", "ce") and mode != "stream" and compress is None and dist.get_backend(group) == "nccl":
            try:
                if backend == "nvl":
                    red = NvlGradReducer(group, num_ctas=int(os.environ.get("ALPRO_NVL_CTAS", "64")))
                else:
                    red = CeGradReducer(group)
                model.engine.grad_alloc = red.alloc
            except Exception as e:             # no symmetric-memory support on this box: NCCL
                import warnings
                warnings.warn(f"alpro_b200: peer-memory gradient reducer unavailable ({type(e).__name__}: {e}); "
                              "falling back to ncclAllReduce")
                red = None
        if red is None:
            red = BucketedAllReduce(group, compress=compress, overlap=(mode != "stream"))
        model._grad_reducer = red
        model.engine.grad_ready_hook = red.ready
    else:
        model._grad_reducer = None
        model.engine.grad_ready_hook = None
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
