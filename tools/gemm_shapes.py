"""Per-shape table of the GEMM launches of ONE training step (event-timed inside the running step, bench batch):
where the tensor-core time goes and which shapes sit furthest below the cuBLAS-sustained peak.
    python tools/gemm_shapes.py [pretrain|retrieval] > gpurun_out/gemm_shapes.md"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "pretrain"
    from alpro_b200 import ops
    dev = torch.device("cuda", 0)
    model = bench.build_model(kind, dev)
    batch = bench.make_batch(kind, 32, 1234, dev)

    def step():
        out = model(batch)
        sum(v for k, v in out.items() if k.endswith("_loss") and v is not None).backward()
        for p in model.parameters():
            p.grad = None
    for _ in range(3):
        step()
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    reps = 3
    for _ in range(reps):
        ops.GEMM_PROFILE = []
        step()
        torch.cuda.synchronize()
        prof, ops.GEMM_PROFILE = ops.GEMM_PROFILE, None
        for M, N, K, e0, e1, nb, tag in prof:
            a = agg[(M, N, K, tag)]
            a[0] += 1
            a[1] += e0.elapsed_time(e1)
            a[2] += 2.0 * M * N * K
            a[3] += nb
    peak = bench.load_peaks()
    tot = sum(a[1] for a in agg.values()) / reps
    print(f"# GEMM launches of one {kind} step (B=32), event-timed in the running step, mean of {reps} steps\n")
    print(f"total {tot:.2f} ms; peak = {peak['sustained']} TFLOP/s (cuBLAS bf16 sustained), HBM {peak['hbm_gbs']} GB/s\n")
    print("| ms/step | launches | M | N | K | mode | TFLOP/s | of peak | GB/s (algorithmic) | ms lost vs peak |")
    print("|---:|---:|---:|---:|---:|---|---:|---:|---:|---:|")
    for (M, N, K, tag), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        ms = a[1] / reps
        tf = a[2] / a[1] / 1e9
        gbs = a[3] / a[1] / 1e6
        ideal = max(a[2] / reps / (peak["sustained"] * 1e9), a[3] / reps / (peak["hbm_gbs"] * 1e6))
        if ms < 0.02:
            continue
        print(f"| {ms:.3f} | {a[0] // reps} | {M} | {N} | {K} | {tag} | {tf:.0f} | {tf / peak['sustained']:.2f} | {gbs:.0f} | {ms - ideal:.3f} |")


if __name__ == "__main__":
    main()
