"""CPU, world_size 2 over gloo: pins the collective semantics of the hot path.

Horovod 0.19.4 is not vendored in the reference, so nothing in the reference pins hvd.allgather's gradient or the
gradient averaging; we pin them here: data-parallel VTC (all-gather forward, reduce-scatter-sum backward, gradient
average) must equal the single-process computation of the mean-over-ranks loss on the concatenated batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Gather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, comm):
        ctx.comm = comm
        return comm.all_gather(x)

    @staticmethod
    def backward(ctx, g):
        return ctx.comm.reduce_scatter_sum(g), None


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from alpro_b200.comm import TorchDistComm
    from oracle import alpro_oracle
    comm = TorchDistComm()
    assert (comm.rank, comm.world) == (rank, world)
    # 1) plain semantics
    x = torch.full((3, 4), float(rank))
    gth = comm.all_gather(x)
    assert gth.shape == (3 * world, 4) and all(float(gth[3 * r, 0]) == r for r in range(world))
    g = torch.arange(3 * world * 4, dtype=torch.float32).view(3 * world, 4) * (rank + 1)
    rs = comm.reduce_scatter_sum(g.clone())
    full = torch.arange(3 * world * 4, dtype=torch.float32).view(3 * world, 4) * sum(r + 1 for r in range(world))
    assert torch.equal(rs, full[3 * rank:3 * rank + 3])
    # 1b) bucketed gradient averaging: prefixes become final in several steps, result = mean over ranks
    from alpro_b200.comm import BucketedAllReduce

    class FakeG:
        pass
    G = FakeG()
    G.flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    red = BucketedAllReduce(min_bucket=300)
    for end in (100, 250, 420, 700, 990):
        red.ready(G, end)
    assert red._done in (420, 990) or red._done >= 300
    nbytes = red.finish()
    assert nbytes == 4000
    want = torch.arange(1000, dtype=torch.float32) * (sum(r + 1 for r in range(world)) / world)
    assert torch.allclose(G.flat, want)
    # 2) VTC through the oracle with a differentiable gather
    b, d = 4, 32
    gen = torch.Generator().manual_seed(5)
    vid_all = torch.randn(world * b, d, generator=gen)
    txt_all = torch.randn(world * b, d, generator=gen)
    sd = {"temp": torch.tensor(0.07), "vision_proj.weight": torch.randn(256, d, generator=gen) * 0.1,
          "vision_proj.bias": torch.zeros(256), "text_proj.weight": torch.randn(256, d, generator=gen) * 0.1,
          "text_proj.bias": torch.zeros(256)}
    for v in sd.values():
        v.requires_grad_(True)
    loss, *_ = alpro_oracle.vtc(sd, "", vid_all[rank * b:(rank + 1) * b], txt_all[rank * b:(rank + 1) * b], rank,
                                gather=lambda t: _Gather.apply(t, comm))
    loss.backward()
    grads = {}
    for k, v in sd.items():                     # gradient averaging (hvd.DistributedOptimizer, op=Average)
        gk = v.grad.clone()
        dist.all_reduce(gk)
        grads[k] = gk / world
    if rank == 0:
        q.put({k: v.numpy() for k, v in grads.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_vtc_data_parallel_equals_global_computation():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process statement of the same objective: mean over ranks of the local VTC losses
    b, d = 4, 32
    gen = torch.Generator().manual_seed(5)
    vid_all = torch.randn(world * b, d, generator=gen)
    txt_all = torch.randn(world * b, d, generator=gen)
    sd = {"temp": torch.tensor(0.07), "vision_proj.weight": torch.randn(256, d, generator=gen) * 0.1,
          "vision_proj.bias": torch.zeros(256), "text_proj.weight": torch.randn(256, d, generator=gen) * 0.1,
          "text_proj.bias": torch.zeros(256)}
    for v in sd.values():
        v.requires_grad_(True)
    vf = F.normalize(F.linear(vid_all, sd["vision_proj.weight"], sd["vision_proj.bias"]), dim=-1)
    tf = F.normalize(F.linear(txt_all, sd["text_proj.weight"], sd["text_proj.bias"]), dim=-1)
    tgt = torch.arange(world * b)
    loss = 0.5 * (F.cross_entropy(vf @ tf.t() / sd["temp"], tgt) + F.cross_entropy(tf @ vf.t() / sd["temp"], tgt))
    loss.backward()
    for k, v in sd.items():
        ref = v.grad.numpy()
        assert abs(got[k] - ref).max() <= 1e-5 * max(1.0, abs(ref).max()), k


def _hvd_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from alpro_b200.shims import hvd
    hvd.init()                                     # joins the group from the torchrun-style environment (gloo: no CUDA)
    ok = (hvd.rank(), hvd.size(), hvd.local_rank()) == (rank, world, rank)
    # allgather: ragged first dimension, rank order (src/utils/distributed.py:234-235 gathers sizes then payloads)
    x = torch.full((rank + 2, 3), float(rank))
    g = hvd.allgather(x)
    ok = ok and g.shape == (sum(r + 2 for r in range(world)), 3) and float(g[0, 0]) == 0.0 and float(g[-1, 0]) == world - 1
    # its gradient: sum over ranks, local slice (Horovod 0.19.4 allgather backward)
    y = torch.ones(2, 4, requires_grad=True)
    (hvd.allgather(y) * (rank + 1)).sum().backward()
    ok = ok and torch.allclose(y.grad, torch.full((2, 4), float(sum(r + 1 for r in range(world)))))
    t = torch.arange(5.0) * (rank + 1)
    hvd.allreduce_(t)                              # average
    ok = ok and torch.allclose(t, torch.arange(5.0) * (sum(r + 1 for r in range(world)) / world))
    b = torch.full((3,), float(rank))
    hvd.broadcast_(b, root_rank=1)
    ok = ok and float(b[0]) == 1.0
    # DistributedOptimizer: gradients averaged by synchronize(), parameters broadcast from rank 0
    torch.manual_seed(rank)
    lin = torch.nn.Linear(4, 2)
    hvd.broadcast_parameters(lin.state_dict(), root_rank=0)
    opt = hvd.DistributedOptimizer(torch.optim.SGD(lin.parameters(), lr=1.0), named_parameters=lin.named_parameters())
    hvd.broadcast_optimizer_state(opt, root_rank=0)
    w0 = lin.weight.detach().clone()
    lin(torch.full((1, 4), float(rank + 1))).sum().backward()
    opt.synchronize()
    want = torch.full((2, 4), sum(r + 1 for r in range(world)) / world)
    ok = ok and torch.allclose(lin.weight.grad, want)
    with opt.skip_synchronize():
        opt.step()
    ok = ok and torch.allclose(lin.weight.detach(), w0 - want)
    q.put((rank, bool(ok), lin.weight.detach().numpy().copy()))
    dist.barrier()
    hvd.shutdown()


def test_horovod_stand_in_world_size_2():
    """alpro_b200.shims.hvd over gloo: the collective semantics the unchanged run scripts rely on."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_hvd_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {r: (ok, w) for r, ok, w in (q.get(timeout=120) for _ in range(world))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][0] and got[1][0]
    assert (got[0][1] == got[1][1]).all()          # same start (broadcast) + same averaged gradient = same weights
