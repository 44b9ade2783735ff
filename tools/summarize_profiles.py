"""Turns gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum) and a full ncu capture into the small text
summaries committed under profiles/. Usage: python tools/summarize_profiles.py <tag> [launches.csv] [rep.ncu-rep]"""
import collections
import csv
import re
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for r in data:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        v = v / 1e3 if r[iu] == "ns" else (v * 1e3 if r[iu] == "ms" else v)
        name = re.sub(r"\(.*", "", r[ik])
        name = re.sub(r"alpro::(\(anonymous namespace\)|<unnamed>)::", "", name)
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    out = [f"total {tot / 1e3:.3f} ms over {len(data)} launches (ncu per-launch times: cold-cache, serialised; compare shares)",
           "", "| ms | share | launches | kernel |", "|---:|---:|---:|---|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if t / tot < 0.0005:
            continue
        out.append(f"| {t / 1e3:.3f} | {100 * t / tot:.1f}% | {n} | `{k[:90]}` |")
    return "\n".join(out)


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "launch__cluster_size", "smsp__cycles_active.avg"]


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        out.append(f"### {d.get('Kernel Name', '?')[:100]}")
        for h in hdr:
            if any(h == w or h.startswith(w) for w in WANT) or "issue_stalled" in h and "per_issue_active" in h:
                try:
                    if float(d[h].replace(",", "")) == 0 and "stalled" in h:
                        continue
                except ValueError:
                    pass
                out.append(f"  {h} = {d[h]}")
    return "\n".join(out)


if __name__ == "__main__":
    tag = sys.argv[1]
    lp = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/launches.csv"
    open(f"profiles/{tag}_launches.md", "w").write(f"# {tag}: kernel launch list of one training step (bench.py --ncu)\n\n" + launches(lp) + "\n")
    if len(sys.argv) > 3:
        open(f"profiles/{tag}_gemm_ncu.txt", "w").write(full(sys.argv[3]) + "\n")
    print(open(f"profiles/{tag}_launches.md").read()[:1500])
