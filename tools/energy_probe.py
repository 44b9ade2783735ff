"""Joules per launch of the step's main kernels (NVML board-energy counter), next to their duration under a sustained
loop. The training step runs AT the board's power limit (bench.py "power"), so its wall time is joules / limit: this
table says which kernels the joules go to, which a duration list cannot. Each case loops one kernel for ~1.2 s (the
counter updates every ~100 ms), buffers are rotated so that operands come from HBM, not L2.
    python tools/energy_probe.py [out.md]          (GPU box; a few tens of seconds)"""
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import pynvml  # noqa: E402

from alpro_b200 import ops  # noqa: E402
from alpro_b200.ops import ACT_GELU, ACT_GELU_GRAD, MNMAJOR  # noqa: E402

OUT = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "energy_probe.md")
LOOP_S = float(os.environ.get("PROBE_SECONDS", "1.2"))
dev = "cuda"
pynvml.nvmlInit()
H = pynvml.nvmlDeviceGetHandleByIndex(0)
LIMIT_W = pynvml.nvmlDeviceGetEnforcedPowerLimit(H) / 1e3


def energy_j():
    return pynvml.nvmlDeviceGetTotalEnergyConsumption(H) / 1e3


def measure(fns):
    """fns: list of closures doing the same work on different buffers (rotated). Returns (us/launch, J/launch, W)."""
    n = len(fns)
    for f in fns:
        f()
    torch.cuda.synchronize()
    # calibrate the launch count for ~LOOP_S seconds
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(8):
        fns[i % n]()
    e1.record()
    torch.cuda.synchronize()
    per = e0.elapsed_time(e1) / 8 * 1e-3
    iters = max(16, int(LOOP_S / per))
    time.sleep(0.15)
    torch.cuda.synchronize()
    j0, t0 = energy_j(), time.perf_counter()
    e0.record()
    for i in range(iters):
        fns[i % n]()
    e1.record()
    torch.cuda.synchronize()
    j1, t1 = energy_j(), time.perf_counter()
    ms = e0.elapsed_time(e1)
    return ms / iters * 1e3, (j1 - j0) / iters, (j1 - j0) / (t1 - t0), pynvml.nvmlDeviceGetClockInfo(H, pynvml.NVML_CLOCK_SM)


M = 50208
d = 768
g = torch.Generator(device=dev).manual_seed(5)


def rnd16(*shape, s=0.5):
    return (torch.randn(*shape, device=dev, generator=g) * s).half()


ROT = 3   # 3 x (77..308 MB) per operand > 126 MB of L2
rows = []


ONLY = [t for t in os.environ.get("PROBE_ONLY", "").split(",") if t]   # substrings of case names; empty = all


def add(name, fns, flop=0.0, bytes_=0.0):
    if ONLY and not any(t in name for t in ONLY):
        return
    us, j, w, mhz = measure(fns)
    rows.append((name, us, j, w, mhz, flop, bytes_))
    print(f"{name:44s} {us:9.1f} us {j * 1e3:9.2f} mJ {w:7.1f} W {mhz:5d} MHz", flush=True)


# ---- idle and pure-HBM references
time.sleep(0.3)
j0, t0 = energy_j(), time.perf_counter()
time.sleep(1.0)
idle_w = (energy_j() - j0) / (time.perf_counter() - t0)
print(f"idle {idle_w:.1f} W, limit {LIMIT_W:.0f} W", flush=True)
src = [torch.empty(M * 3072, device=dev, dtype=torch.float16) for _ in range(ROT)]
dst = [torch.empty(M * 3072, device=dev, dtype=torch.float16) for _ in range(ROT)]
add("torch copy 308 MB -> 308 MB", [lambda i=i: dst[i].copy_(src[i]) for i in range(ROT)], 0, 2 * M * 3072 * 2)
del src, dst

# ---- GEMMs: ours and cuBLAS at the same shapes
a768 = [rnd16(M, d) for _ in range(ROT)]
a3072 = [rnd16(M, 4 * d, s=0.1) for _ in range(ROT)]
w_qkv, w_fc1, w_fc2, w_sq = rnd16(3 * d, d, s=0.03), rnd16(4 * d, d, s=0.03), rnd16(d, 4 * d, s=0.03), rnd16(d, d, s=0.03)
b2304, b3072, b768 = (torch.randn(n, device=dev, generator=g) for n in (3 * d, 4 * d, d))
o2304 = [torch.empty(M, 3 * d, device=dev, dtype=torch.float16) for _ in range(ROT)]
o3072 = [torch.empty(M, 4 * d, device=dev, dtype=torch.float16) for _ in range(ROT)]
o3072b = [torch.empty(M, 4 * d, device=dev, dtype=torch.float16) for _ in range(ROT)]
o768 = [torch.empty(M, d, device=dev, dtype=torch.float16) for _ in range(ROT)]
x32 = [torch.randn(M, d, device=dev, generator=g) for _ in range(ROT)]
y32 = [torch.empty(M, d, device=dev) for _ in range(ROT)]

add("ours  qkv  50208x2304x768 +bias o16", [lambda i=i: ops.gemm16(a768[i], w_qkv, bias=b2304, out16=o2304[i])
                                             for i in range(ROT)], 2.0 * M * 2304 * 768, M * (768 + 2304) * 2)
add("cuBLAS     50208x2304x768 (addmm)", [lambda i=i: torch.matmul(a768[i], w_qkv.t(), out=o2304[i])
                                           for i in range(ROT)], 2.0 * M * 2304 * 768, M * (768 + 2304) * 2)
add("ours  fc1  50208x3072x768 +GELU o16+o16b",
    [lambda i=i: ops.gemm16(a768[i], w_fc1, bias=b3072, act=ACT_GELU, out16=o3072[i], out16b=o3072b[i])
     for i in range(ROT)], 2.0 * M * 3072 * 768, M * (768 + 2 * 3072) * 2)
add("ours  fc1  50208x3072x768 +GELU o16 (eval)",
    [lambda i=i: ops.gemm16(a768[i], w_fc1, bias=b3072, act=ACT_GELU, out16=o3072[i]) for i in range(ROT)],
    2.0 * M * 3072 * 768, M * (768 + 3072) * 2)
add("cuBLAS     50208x3072x768", [lambda i=i: torch.matmul(a768[i], w_fc1.t(), out=o3072[i]) for i in range(ROT)],
    2.0 * M * 3072 * 768, M * (768 + 3072) * 2)
add("ours  fc2  50208x768x3072 +resid o32",
    [lambda i=i: ops.gemm16(a3072[i], w_fc2, bias=b768, resid=x32[i], out32=y32[i]) for i in range(ROT)],
    2.0 * M * 768 * 3072, M * (3072 * 2 + 768 * 8))
add("cuBLAS     50208x768x3072", [lambda i=i: torch.matmul(a3072[i], w_fc2.t(), out=o768[i]) for i in range(ROT)],
    2.0 * M * 768 * 3072, M * (3072 + 768) * 2)
add("ours  proj 50208x768x768 +resid o32",
    [lambda i=i: ops.gemm16(a768[i], w_sq, bias=b768, resid=x32[i], out32=y32[i]) for i in range(ROT)],
    2.0 * M * 768 * 768, M * (768 * 2 + 768 * 8))
add("cuBLAS     50208x768x768", [lambda i=i: torch.matmul(a768[i], w_sq.t(), out=o768[i]) for i in range(ROT)],
    2.0 * M * 768 * 768, M * (768 + 768) * 2)
add("ours  dgrad fc2 50208x3072x768 *gelu' o16",
    [lambda i=i: ops.gemm16(a768[i], w_fc2, b_layout=MNMAJOR, act=ACT_GELU_GRAD, aux=o3072b[i], out16=o3072[i])
     for i in range(ROT)], 2.0 * M * 3072 * 768, M * (768 + 2 * 3072) * 2)
add("ours  dgrad fc1 50208x768x3072 o16",
    [lambda i=i: ops.gemm16(a3072[i], w_fc1, b_layout=MNMAJOR, out16=o768[i]) for i in range(ROT)],
    2.0 * M * 768 * 3072, M * (3072 + 768) * 2)
gw = torch.zeros(4 * d, d, device=dev)
add("ours  wgrad fc1 3072x768x50208 split-K",
    [lambda i=i: ops.gemm16(a3072[i], a768[i], a_layout=MNMAJOR, b_layout=MNMAJOR, out32=gw, split_k=-1)
     for i in range(ROT)], 2.0 * M * 768 * 3072, M * (3072 + 768) * 2)
gw2 = torch.zeros(d, d, device=dev)
add("ours  wgrad proj 768x768x50208 split-K",
    [lambda i=i: ops.gemm16(a768[i], a768[(i + 1) % ROT], a_layout=MNMAJOR, b_layout=MNMAJOR, out32=gw2, split_k=-1)
     for i in range(ROT)], 2.0 * M * 768 * 768, M * (768 + 768) * 2)
big = [rnd16(8192, 8192, s=0.05) for _ in range(2)]
bo = torch.empty(8192, 8192, device=dev, dtype=torch.float16)
add("cuBLAS     8192^3", [lambda: torch.matmul(big[0], big[1], out=bo)], 2.0 * 8192 ** 3, 3 * 8192 * 8192 * 2)
add("ours       8192^3", [lambda: ops.gemm16(big[0], big[1], out16=bo)], 2.0 * 8192 ** 3, 3 * 8192 * 8192 * 2)
del big, bo

# ---- LayerNorm, column sums
gamma, beta = torch.ones(d, device=dev), torch.zeros(d, device=dev)
st = torch.empty(2, M, device=dev)
add("layernorm_fwd 50208x768 (x32 -> a16)",
    [lambda i=i: ops.layernorm_fwd(x32[i], gamma, beta, 1e-6, out16=o768[i], mean=st[0], rstd=st[1]) for i in range(ROT)],
    0, M * d * 6)
dgam, dbet, cs = torch.zeros(d, device=dev), torch.zeros(d, device=dev), torch.zeros(d, device=dev)
add("layernorm_bwd 50208x768 (da16,x32,dx32+= ,dx16)",
    [lambda i=i: ops.layernorm_bwd(a768[i], x32[i], st[0], st[1], gamma, y32[i], 1, dx16=o768[i], dgamma=dgam, dbeta=dbet,
                                   param_scale=1.0, colsum=cs) for i in range(ROT)], 0, M * d * 16)
cs3072 = torch.zeros(4 * d, device=dev)
add("colsum16 50208x3072", [lambda i=i: ops.colsum(a3072[i], cs3072, 1.0, 0) for i in range(ROT)], 0, M * 3072 * 2)
cs2304 = torch.zeros(3 * d, device=dev)
add("colsum16 50208x2304", [lambda i=i: ops.colsum(o2304[i], cs2304, 1.0, 0) for i in range(ROT)], 0, M * 2304 * 2)

# ---- attention at the ViT shape
B, T, N, heads = 32, 8, 196, 12
Sc, S, nseq = 1 + N * T, 1 + N, B * T
qkv = [rnd16(M, 3 * d) for _ in range(ROT)]
cls_o = torch.empty(nseq, d, device=dev, dtype=torch.float16)
lse = torch.empty(nseq, heads, S, device=dev)
scratch = torch.empty(nseq, 3 * d, device=dev)
add("sattn fwd 256x12 S=197",
    [lambda i=i: ops.seq_attn_fwd(qkv[i], None, o768[i], cls_o, lse, S, nseq, heads, T, T, Sc, 0.125) for i in range(ROT)],
    4.0 * S * S * 64 * nseq * heads, M * d * 2 * 4)
add("sattn bwd 256x12 S=197",
    [lambda i=i: ops.seq_attn_bwd(qkv[i], None, lse, o768[i], cls_o, a768[i], o2304[i], scratch, S, nseq, heads, T, T, Sc,
                                  0.125) for i in range(ROT)], 10.0 * S * S * 64 * nseq * heads, M * d * 2 * 8)
add("tattn fwd 6272x12 T=8", [lambda i=i: ops.temporal_attn_fwd(qkv[i], o768[i], B, N, T, heads, 0.125)
                              for i in range(ROT)], 0, M * d * 2 * 4)
add("tattn bwd 6272x12 T=8", [lambda i=i: ops.temporal_attn_bwd(qkv[i], a768[i], o2304[i], B, N, T, heads, 0.125)
                              for i in range(ROT)], 0, M * d * 2 * 7)

lines = ["# joules per launch (NVML energy counter), one kernel looped ~%.1f s, operands rotated through HBM" % LOOP_S, "",
         f"board idle {idle_w:.0f} W, enforced power limit {LIMIT_W:.0f} W. `dyn mJ` = energy above idle for the kernel's "
         "duration; pJ/FLOP and pJ/B use the dynamic part.", "",
         "| kernel | us | mJ | W | SM MHz | dyn mJ | TFLOP/s | pJ/FLOP (dyn) | GB/s (algorithmic) | pJ/B (dyn) |",
         "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
for name, us, j, w, mhz, flop, by in rows:
    dyn = j - idle_w * us * 1e-6
    lines.append(f"| {name} | {us:.1f} | {j * 1e3:.2f} | {w:.0f} | {mhz} | {dyn * 1e3:.2f} | "
                 f"{(flop / us * 1e-6) if flop else 0:.0f} | {(dyn / flop * 1e12) if flop else 0:.3f} | "
                 f"{(by / us * 1e-3) if by else 0:.0f} | {(dyn / by * 1e12) if by else 0:.1f} |")
os.makedirs(os.path.dirname(OUT), exist_ok=True)
open(OUT, "w").write("\n".join(lines) + "\n")
