"""Opcode mix of one profiled kernel launch: instructions executed and stall samples per SASS opcode.
   python tools/ncu_opmix.py rep.ncu-rep [launch_index] [top_n]   (needs --import-source on / --set full captures)"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 28
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(skip),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
name = rows[0][1][:120] if rows and len(rows[0]) > 1 else "?"
hdr = None
ops, samp = collections.Counter(), collections.Counter()
lines = []
tot = 0
for r in rows:
    if "Instructions Executed" in r and "Source" in r:
        hdr = r
        ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= ie:
        continue
    try:
        n = int(r[ie])
    except ValueError:
        continue
    toks = r[ia].strip().split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "LDG", "STG", "MUFU", "LDGSTS", "ATOM", "RED", "BAR", "SHFL")) else op.split(".")[0]
    ops[op] += n
    samp[op] += int(r[isamp])
    tot += n
    lines.append((int(r[isamp]), n, r[ia].strip()))
print(name)
print("total warp instructions", tot, " total samples", sum(samp.values()))
for k, v in ops.most_common(top):
    print(f"{k:14s} {v:11d} {100 * v / tot:5.1f}%   samples {samp[k]:7d} {100 * samp[k] / max(1, sum(samp.values())):5.1f}%")
print("-- hottest instructions by stall samples")
for s, n, src in sorted(lines, reverse=True)[:14]:
    print(f"{s:7d} {n:10d}  {src[:100]}")
