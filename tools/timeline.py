"""Device timeline of one training step from CUPTI (torch.profiler): busy time, idle gaps between kernels, and which
kernel pairs the gaps sit between. Answers "is the step bubbles or kernels?" with the in-step (power-capped, warm)
durations that ncu's serialised replay cannot give. Run on the GPU box:
    python tools/timeline.py [batch] [out.md]
A number printed here is taken under the profiler: it explains the bench value, it is never one."""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
OUT = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "timeline.md")
STEPS = 3
dev = torch.device("cuda", 0)
model = bench.build_model("pretrain", dev)
model.train()
batch = bench.make_batch("pretrain", B, 1234, dev)


def step():
    out = model(batch)
    sum(v for k, v in out.items() if k.endswith("_loss") and v is not None).backward()
    for p in model.parameters():
        p.grad = None


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(STEPS):
    step()
e1.record()
torch.cuda.synchronize()
plain_ms = e0.elapsed_time(e1) / STEPS

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(STEPS):
        step()
    torch.cuda.synchronize()

ev = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start:
        ev.append((e.time_range.start, e.time_range.end, e.name))
ev.sort()
if not ev:
    raise SystemExit("no CUDA events captured")
t_first, t_last = ev[0][0], max(e[1] for e in ev)
span_ms = (t_last - t_first) / 1e3 / STEPS


def short(n):
    for pre in ("(anonymous namespace)::", "void ", "alpro::", "gemm3::", "gemm2::", "at::native::", "at::"):
        n = n.replace(pre, "")
    n = n.split("(")[0]
    return n[:64]


busy = collections.Counter()
count = collections.Counter()
gap_after = collections.Counter()
gap_pairs = collections.Counter()
gap_hist = collections.Counter()
cursor = ev[0][0]
prev = None
idle = 0.0
for s, e, n in ev:
    n = short(n)
    if s > cursor:
        g = s - cursor
        idle += g
        if prev is not None:
            gap_after[prev] += g
            gap_pairs[(prev, n)] += g
        gap_hist[min(int(g), 20)] += 1
    busy[n] += e - s
    count[n] += 1
    if e > cursor:
        cursor = e
        prev = n

lines = [f"# device timeline of the pretrain step (B={B}), CUPTI via torch.profiler, {STEPS} steps averaged", ""]
lines.append(f"step without profiler {plain_ms:.3f} ms; span under profiler {span_ms:.3f} ms/step; "
             f"kernel-busy {sum(busy.values()) / 1e3 / STEPS:.3f} ms/step (sum of durations, overlapping streams counted "
             f"twice); idle gaps on the device {idle / 1e3 / STEPS:.3f} ms/step over {len(ev) // STEPS} launches")
lines += ["", "## busy time by kernel (in-step durations)", "", "| ms/step | launches/step | mean us | kernel |", "|---:|---:|---:|---|"]
for n, t in busy.most_common(40):
    lines.append(f"| {t / 1e3 / STEPS:.3f} | {count[n] / STEPS:.0f} | {t / count[n]:.1f} | `{n}` |")
lines += ["", "## idle time by the kernel BEFORE the gap", "", "| ms/step | kernel |", "|---:|---|"]
for n, t in gap_after.most_common(25):
    lines.append(f"| {t / 1e3 / STEPS:.3f} | `{n}` |")
lines += ["", "## largest gap pairs", "", "| ms/step | before -> after |", "|---:|---|"]
for (a, b), t in gap_pairs.most_common(25):
    lines.append(f"| {t / 1e3 / STEPS:.3f} | `{a}` -> `{b}` |")
lines += ["", "## gap length histogram (us, last bin = 20+)", ""]
lines.append(" ".join(f"{k}:{gap_hist[k] // STEPS}" for k in sorted(gap_hist)))
os.makedirs(os.path.dirname(OUT), exist_ok=True)
open(OUT, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:3]))
