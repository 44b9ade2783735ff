"""Names the data-parallel scaling limiter with numbers. Run under torchrun on N GPUs, once per
variant (the variant's knobs are environment variables read at attach time / by NCCL):

    torchrun --nproc-per-node 2 tools/scaling_probe.py <tag>

Prints one JSON line: step time (max over ranks), the event-timed sum of all GEMM launches of one step (GEMMs that run
next to a concurrent NCCL kernel show up here), backward wall time, and the exposed tail (allreduce_gradients call)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "probe"
    D = bench.Dist()
    from alpro_b200 import comm as acomm, ops
    kind = os.environ.get("PROBE_KIND", "pretrain")
    model = bench.build_model(kind, D.device)
    use_comm = os.environ.get("PROBE_COMM", "1") == "1"
    if D.world > 1 and use_comm:
        acomm.attach(model)
    skip_grad_reduce = os.environ.get("PROBE_NO_GRAD_REDUCE", "0") == "1"
    if skip_grad_reduce:
        model._grad_reducer = None
        model.engine.grad_ready_hook = None
    batch = bench.make_batch(kind, 32, 1234 + D.rank, D.device)
    marks = {}

    def step(timed=False):
        out = model(batch)
        loss = sum(v for k, v in out.items() if k.endswith("_loss") and v is not None)
        if timed:
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
        loss.backward()
        if timed:
            e[1].record()
        if D.world > 1 and use_comm and not skip_grad_reduce:
            acomm.allreduce_gradients(model)
        if timed:
            e[2].record()
            marks["e"] = e
        for p in model.parameters():
            p.grad = None

    for _ in range(3):
        step()
    ms = D.timed(step, 6)
    step(timed=True)
    torch.cuda.synchronize()
    e = marks["e"]
    bwd_ms, tail_ms = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
    ops.GEMM_PROFILE = []
    step()
    torch.cuda.synchronize()
    prof, ops.GEMM_PROFILE = ops.GEMM_PROFILE, None
    gemm_ms = sum(r[3].elapsed_time(r[4]) for r in prof)
    # GEMMs of the backward pass only (the ones that can overlap a bucket reduction): the last 2/3 of the launches
    worst = sorted((r[3].elapsed_time(r[4]), r[0], r[1], r[2]) for r in prof)[-5:]
    t = torch.tensor([bwd_ms, tail_ms, gemm_ms], device=D.device)
    if D.world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if D.rank == 0:
        print(json.dumps({"tag": tag, "world": D.world, "ms_per_step": round(ms, 3), "pairs_per_s": round(32 * D.world / ms * 1e3, 1),
                          "bwd_ms": round(float(t[0]), 3), "exposed_allreduce_ms": round(float(t[1]), 3),
                          "gemm_ms_sum": round(float(t[2]), 3), "slowest_gemms_ms": [[round(w[0], 3)] + list(w[1:]) for w in worst],
                          "env": {k: os.environ[k] for k in os.environ if k.startswith(("ALPRO_", "NCCL_", "PROBE_"))}}),
              flush=True)
    D.close()


if __name__ == "__main__":
    main()
