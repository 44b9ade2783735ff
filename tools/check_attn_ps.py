#!/usr/bin/env python
"""GPU check of the persistent sequence-attention forward (sattn_ps.cu, ALPRO_ATTN_PS=1) against the per-unit tcgen05
kernel and a torch fp32 reference; every case in its own subprocess under a timeout (a hung pipeline must not take the
box down). One JSON line per case on stdout and in gpurun_out/attn_ps_check.jsonl.
    python tools/check_attn_ps.py            # all cases
    python tools/check_attn_ps.py --case 2:8:196:3:fp16"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CASES = ["1:2:196:1:fp16", "2:2:196:3:fp16", "2:8:196:3:fp16", "1:4:160:2:fp16", "1:2:255:2:fp16", "1:3:128:2:bf16",
         "3:8:196:12:bf16", "time:32:8:196:12"]


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def run(case):
    import torch
    from alpro_b200 import ops
    timed = case.startswith("time:")
    parts = case.split(":")
    if timed:
        B, T, N, heads = (int(x) for x in parts[1:5])
        dt = torch.float16
    else:
        B, T, N, heads = (int(x) for x in parts[:4])
        dt = torch.float16 if parts[4] == "fp16" else torch.bfloat16
    dev = "cuda"
    d = heads * 64
    Sc, S, nseq = 1 + N * T, 1 + N, B * T
    gen = torch.Generator(device=dev).manual_seed(13)
    qkv = (torch.randn(B * Sc, 3 * d, device=dev, generator=gen)).to(dt)
    outs = {}
    res = {"case": case, "S": S, "nseq": nseq, "heads": heads}
    flush = torch.empty(160 * 1024 * 1024, device=dev, dtype=torch.uint8)
    for name, flag in (("tc", "0"), ("ps", "1")):
        os.environ["ALPRO_ATTN_PS"] = flag
        o = torch.full((B * Sc, d), 7.0, device=dev, dtype=dt)
        cls_o = torch.full((nseq, d), 7.0, device=dev, dtype=dt)
        lse = torch.full((nseq, heads, S), 7.0, device=dev)
        ops.seq_attn_fwd(qkv, None, o, cls_o, lse, S, nseq, heads, T, T, Sc, 0.125)
        torch.cuda.synchronize()
        outs[name] = (o.float().clone(), cls_o.float().clone(), lse.clone())
        if timed:
            for _ in range(3):
                ops.seq_attn_fwd(qkv, None, o, cls_o, lse, S, nseq, heads, T, T, Sc, 0.125)
            torch.cuda.synchronize()
            tot = 0.0
            for _ in range(8):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.seq_attn_fwd(qkv, None, o, cls_o, lse, S, nseq, heads, T, T, Sc, 0.125)
                e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            res[f"{name}_ms"] = round(tot / 8, 4)
    # rows of o: the clip's cls row is written by cls_mean_fwd later (both kernels leave it alone): compare patch rows
    ov = {k: v[0].view(B, Sc, d)[:, 1:] for k, v in outs.items()}
    res["o_ps_vs_tc"] = rel(ov["ps"], ov["tc"])
    res["cls_ps_vs_tc"] = rel(outs["ps"][1], outs["tc"][1])
    res["lse_ps_vs_tc"] = float((outs["ps"][2] - outs["tc"][2]).abs().max())
    res["cls_row_untouched"] = bool((outs["ps"][0].view(B, Sc, d)[:, 0] == 7.0).all())
    res["finite"] = bool(torch.isfinite(outs["ps"][0]).all() and torch.isfinite(outs["ps"][2]).all())
    if not timed or B <= 4:
        x = qkv.float().view(B, Sc, 3 * d)
        cls = x[:, :1].unsqueeze(1).expand(B, T, 1, 3 * d)
        pat = x[:, 1:].view(B, N, T, 3 * d).permute(0, 2, 1, 3)
        xs = torch.cat([cls, pat], 2).reshape(B * T, S, 3, heads, 64).permute(2, 0, 3, 1, 4)
        sc = (xs[0] @ xs[1].transpose(-1, -2)) * 0.125
        ref = (torch.softmax(sc, -1) @ xs[2]).permute(0, 2, 1, 3).reshape(B, T, S, d)
        want = ref[:, :, 1:].permute(0, 2, 1, 3).reshape(B, N * T, d)
        res["o_ps_vs_ref"] = rel(ov["ps"], want)
        res["o_tc_vs_ref"] = rel(ov["tc"], want)
        res["cls_ps_vs_ref"] = rel(outs["ps"][1].view(B, T, d), ref[:, :, 0])
        lse_ref = torch.logsumexp(sc, -1) * 1.4426950408889634   # [nseq, heads, S] base 2
        res["lse_ps_vs_ref"] = float((outs["ps"][2] - lse_ref).abs().max())
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case")
    args = ap.parse_args()
    if args.case:
        print(json.dumps(run(args.case)), flush=True)
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    bad = 0
    with open(os.path.join(ROOT, "gpurun_out", "attn_ps_check.jsonl"), "w") as f:
        for c in CASES:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", c], capture_output=True, text=True,
                                   timeout=150)
                line = [l for l in r.stdout.splitlines() if l.startswith("{")]
                out = line[-1] if line else json.dumps({"case": c, "error": (r.stderr or r.stdout)[-400:]})
            except subprocess.TimeoutExpired:
                out = json.dumps({"case": c, "error": "timeout (hung kernel?)"})
            print(out, flush=True)
            f.write(out + "\n")
            d = json.loads(out)
            if "error" in d or not d.get("finite", False) or d.get("o_ps_vs_tc", 1) > 2e-3:
                bad += 1
                if "timeout" in d.get("error", ""):
                    break   # the device may be wedged: stop here
    print(f"check_attn_ps: {bad} problem(s)")


if __name__ == "__main__":
    main()
