#!/usr/bin/env python
"""GPU diagnostic for the tcgen05 sequence-attention kernels (ALPRO_ATTN_TC / ALPRO_ATTN_BWD_TC): every case runs in its
own subprocess under a timeout (a hung kernel must not take the box down), compares the tcgen05 kernel with the
mma.sync kernel of the same library and with a torch fp32 reference, per output block (dq / dk / dv), and times both at
the bench shapes. Results: one JSON line per case on stdout and in gpurun_out/attn_tc_check.jsonl.

  python tools/check_attn_tc.py            # all cases
  python tools/check_attn_tc.py --cases bert:197:fp16:0,vit:196:8      # the listed cases, in-process
"""
import argparse
import json
import math
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = ["bert:16:fp16:0", "bert:40:fp16:0", "bert:64:fp16:0", "bert:128:fp16:0", "bert:129:bf16:0", "bert:144:fp16:0",
         "bert:197:fp16:0", "bert:237:fp16:0", "bert:240:fp16:0", "bert:237:fp16:0.1", "bert:40:fp16:0.1",
         "vit:4:2", "vit:196:2", "vit:9:4", "vit:196:8",
         "time:vit", "time:fusion", "time:text"]


def rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12))


def setenv(fwd_tc, bwd_tc):
    os.environ["ALPRO_ATTN_TC"] = "1" if fwd_tc else "0"
    os.environ["ALPRO_ATTN_BWD_TC"] = "1" if bwd_tc else "0"


def blocks(dqkv, d):
    return {"dq": dqkv[:, :d], "dk": dqkv[:, d:2 * d], "dv": dqkv[:, 2 * d:]}


def run_bert(S, dt, drop):
    import torch
    from alpro_b200 import ops
    dev = "cuda"
    dt = torch.float16 if dt == "fp16" else torch.bfloat16
    nseq, heads = 3, 3
    d = heads * 64
    gen = torch.Generator(device=dev).manual_seed(8)
    qkv = torch.randn(nseq * S, 3 * d, device=dev, generator=gen).to(dt)
    keep = torch.rand(nseq, S, device=dev, generator=gen) > 0.25
    keep[:, 0] = True
    mask = (1.0 - keep.float()) * -10000.0
    scale = 1 / math.sqrt(64)
    o = torch.empty(nseq * S, d, device=dev, dtype=dt)
    lse = torch.empty(nseq, heads, S, device=dev)
    setenv(False, False)
    ops.seq_attn_fwd(qkv, mask, o, None, lse, S, nseq, heads, 1, 1, S, scale, drop, 77)
    do = torch.randn(nseq * S, d, device=dev, generator=gen).to(dt)
    out = {}
    res = {}
    for name, tc in (("mma", False), ("tc", True)):
        setenv(False, tc)
        dqkv = torch.full((nseq * S, 3 * d), 7.0, device=dev, dtype=dt)
        ops.seq_attn_bwd(qkv, mask, lse, o, None, do, dqkv, None, S, nseq, heads, 1, 1, S, scale, None, drop, 77)
        torch.cuda.synchronize()
        out[name] = dqkv.float()
    for k in ("dq", "dk", "dv"):
        res["tc_vs_mma_" + k] = rel(blocks(out["tc"], d)[k], blocks(out["mma"], d)[k])
    res["finite"] = bool(torch.isfinite(out["tc"]).all())
    if drop == 0:
        x = qkv.float().view(nseq, S, 3, heads, 64).permute(2, 0, 3, 1, 4).detach().requires_grad_(True)
        s = (x[0] @ x[1].transpose(-1, -2)) * scale + mask[:, None, None, :]
        ref = torch.softmax(s, -1) @ x[2]
        ref.backward(do.float().view(nseq, S, heads, 64).permute(0, 2, 1, 3))
        want = x.grad.permute(1, 3, 0, 2, 4).reshape(nseq * S, 3 * d)
        for k in ("dq", "dk", "dv"):
            res["tc_vs_ref_" + k] = rel(blocks(out["tc"], d)[k], blocks(want, d)[k])
            res["mma_vs_ref_" + k] = rel(blocks(out["mma"], d)[k], blocks(want, d)[k])
    # forward tcgen05 kernel vs the mma.sync one
    setenv(True, False)
    o2 = torch.empty_like(o)
    lse2 = torch.empty_like(lse)
    ops.seq_attn_fwd(qkv, mask, o2, None, lse2, S, nseq, heads, 1, 1, S, scale, drop, 77)
    torch.cuda.synchronize()
    res["fwd_tc_vs_mma_o"] = rel(o2.float(), o.float())
    res["fwd_tc_vs_mma_lse"] = float((lse2 - lse).abs().max())
    return res


def run_vit(N, T):
    import torch
    from alpro_b200 import ops
    dev = "cuda"
    B, heads = 2, 3
    d = heads * 64
    Sc = 1 + N * T
    S = 1 + N
    gen = torch.Generator(device=dev).manual_seed(9)
    qkv = torch.randn(B * Sc, 3 * d, device=dev, generator=gen).half()
    o = torch.zeros(B * Sc, d, device=dev, dtype=torch.float16)
    cls_o = torch.empty(B * T, d, device=dev, dtype=torch.float16)
    lse = torch.empty(B * T, heads, S, device=dev)
    setenv(False, False)
    ops.seq_attn_fwd(qkv, None, o, cls_o, lse, S, B * T, heads, T, T, Sc, 0.125)
    ops.cls_mean_fwd(cls_o, o, B, T, Sc, d)
    do = torch.randn(B * Sc, d, device=dev, generator=gen).half()
    out = {}
    for name, tc in (("mma", False), ("tc", True)):
        setenv(False, tc)
        dqkv = torch.full((B * Sc, 3 * d), 7.0, device=dev, dtype=torch.float16)
        scratch = torch.empty(B * T, 3 * d, device=dev)
        ops.seq_attn_bwd(qkv, None, lse, o, cls_o, do, dqkv, scratch, S, B * T, heads, T, T, Sc, 0.125)
        torch.cuda.synchronize()
        out[name] = dqkv.float()
    res = {}
    for k in ("dq", "dk", "dv"):
        res["tc_vs_mma_" + k] = rel(blocks(out["tc"], d)[k], blocks(out["mma"], d)[k])
        a, b = blocks(out["tc"], d)[k].view(B, Sc, d), blocks(out["mma"], d)[k].view(B, Sc, d)
        res["tc_vs_mma_cls_" + k] = rel(a[:, 0], b[:, 0])
    res["finite"] = bool(torch.isfinite(out["tc"]).all())
    x = qkv.float().view(B, Sc, 3 * d).detach().requires_grad_(True)
    cls = x[:, :1].unsqueeze(1).expand(B, T, 1, 3 * d)
    pat = x[:, 1:].view(B, N, T, 3 * d).permute(0, 2, 1, 3)
    xs = torch.cat([cls, pat], 2).reshape(B * T, S, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = (torch.softmax((xs[0] @ xs[1].transpose(-1, -2)) * 0.125, -1) @ xs[2]).permute(0, 2, 1, 3).reshape(B, T, S, d)
    want = torch.cat([ref[:, :, 0].mean(1, keepdim=True), ref[:, :, 1:].permute(0, 2, 1, 3).reshape(B, N * T, d)], 1)
    want.backward(do.float().view(B, Sc, d))
    g = x.grad.view(B * Sc, 3 * d)
    for k in ("dq", "dk", "dv"):
        res["tc_vs_ref_" + k] = rel(blocks(out["tc"], d)[k], blocks(g, d)[k])
    return res


def run_time(which, order=(("mma", False, False), ("tc", True, True))):
    import torch
    from alpro_b200 import ops
    dev = "cuda"
    heads, d = 12, 768
    gen = torch.Generator(device=dev).manual_seed(3)
    if which == "vit":
        B, T, N = 32, 8, 196
        Sc, S, nseq, seq_div, stride = 1 + N * T, 1 + N, B * T, T, T
        rows = B * Sc
        mask, drop, scale = None, 0.0, 0.125
    else:
        S = 237 if which == "fusion" else 40
        nseq = 128 if which == "fusion" else 64
        Sc, seq_div, stride = S, 1, 1
        rows = nseq * S
        keep = torch.rand(nseq, S, device=dev, generator=gen) > 0.1
        keep[:, 0] = True
        mask, drop, scale = (1.0 - keep.float()) * -10000.0, 0.1, 0.125
    qkv = (0.5 * torch.randn(rows, 3 * d, device=dev, generator=gen)).half()
    o = torch.zeros(rows, d, device=dev, dtype=torch.float16)
    cls_o = torch.empty(nseq, d, device=dev, dtype=torch.float16) if seq_div > 1 else None
    lse = torch.empty(nseq, heads, S, device=dev)
    do = torch.randn(rows, d, device=dev, generator=gen).half()
    dqkv = torch.empty(rows, 3 * d, device=dev, dtype=torch.float16)
    scratch = torch.empty(nseq, 3 * d, device=dev) if seq_div > 1 else None
    flush = torch.empty(160 * 1024 * 1024, device=dev, dtype=torch.uint8)   # > L2
    res = {"shape": f"nseq={nseq} S={S} heads={heads} seq_div={seq_div} drop={drop}"}

    def timeit(fn, iters=8):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / iters

    def fwd():
        ops.seq_attn_fwd(qkv, mask, o, cls_o, lse, S, nseq, heads, seq_div, stride, Sc, scale, drop, 5)

    def bwd():
        ops.seq_attn_bwd(qkv, mask, lse, o, cls_o, do, dqkv, scratch, S, nseq, heads, seq_div, stride, Sc, scale, None,
                         drop, 5)

    outs = {}
    for name, ftc, btc in order:
        setenv(ftc, btc)
        if name == "tcf":   # forward only, last: leaves the forward kernel's stamps in the trace buffer
            timeit(fwd, 2)
            continue
        res[f"fwd_{name}_ms"] = round(timeit(fwd), 4)
        res[f"bwd_{name}_ms"] = round(timeit(bwd), 4)
        outs[name] = (o.float().clone(), dqkv.float().clone())
    res["fwd_tc_vs_mma"] = rel(outs["tc"][0], outs["mma"][0])
    res["bwd_tc_vs_mma"] = rel(outs["tc"][1], outs["mma"][1])
    flops_bwd = 5 * 2.0 * S * S * 64 * nseq * heads
    res["bwd_tc_tflops"] = round(flops_bwd / (res["bwd_tc_ms"] * 1e-3) / 1e12, 1)
    res["bwd_mma_tflops"] = round(flops_bwd / (res["bwd_mma_ms"] * 1e-3) / 1e12, 1)
    return res


SLOTS = {1: "loads issued", 2: "cp.async landed", 3: "prologue done"}
for _s in range(8):
    SLOTS[4 + 4 * _s] = f"s{_s} S/dP ready"
    SLOTS[5 + 4 * _s] = f"s{_s} tmem drained"
    SLOTS[6 + 4 * _s] = f"s{_s} staging free"
    SLOTS[7 + 4 * _s] = f"s{_s} P/dS staged"
    SLOTS[48 + 2 * _s] = f"  mma: s{_s} p_ready seen"
    SLOTS[49 + 2 * _s] = f"  mma: s{_s} grads issued"
SLOTS.update({36: "dkv0 wait", 37: "dkv0 acc_full", 38: "dkv0 stored", 39: "dkv1 wait", 40: "dkv1 acc_full",
              41: "dkv1 stored", 42: "dq stored", 43: "cta done"})


FSLOTS = {1: "loads issued", 2: "tiles landed"}
for _t in range(2):
    FSLOTS.update({4 + 10 * _t: f"t{_t} S ready", 5 + 10 * _t: f"t{_t} max done", 6 + 10 * _t: f"t{_t} P half0 written",
                   7 + 10 * _t: f"t{_t} PV half0 done", 8 + 10 * _t: f"t{_t} P half1 written",
                   9 + 10 * _t: f"t{_t} PV half1 done", 10 + 10 * _t: f"t{_t} O stored", 11 + 10 * _t: f"t{_t} tile end"})


def run_ftrace(which):
    """Per-phase timeline of the tcgen05 forward kernel."""
    import ctypes
    import torch
    os.environ["ALPRO_ATTN_TRACE"] = "1"
    res = run_time(which, order=(("tc", True, True), ("mma", False, False), ("tcf", True, False)))
    from alpro_b200 import _lib
    n = 12 * 256 * 64
    buf = (ctypes.c_int64 * n)()
    nct = _lib.lib.alpro_debug_attn_trace(ctypes.cast(buf, ctypes.c_void_p), n)
    t = torch.tensor(list(buf[: nct * 64]), dtype=torch.float64).view(nct, 64)
    med = t.median(0).values
    order = sorted((float(med[k]), k) for k in FSLOTS if float(med[k]) > 0)
    res["timeline"] = [f"{int(v):6d} {FSLOTS[k]}" for v, k in order]
    res["ctas"] = nct
    return res


def run_trace(which):
    """Per-phase timeline of the tcgen05 backward kernel (median over CTAs of the clock64 stamps, in cycles)."""
    import ctypes
    import torch
    os.environ["ALPRO_ATTN_TRACE"] = "1"
    res = run_time(which)
    from alpro_b200 import _lib
    n = 12 * 256 * 64
    buf = (ctypes.c_int64 * n)()
    nct = _lib.lib.alpro_debug_attn_trace(ctypes.cast(buf, ctypes.c_void_p), n)
    t = torch.tensor(list(buf[: nct * 64]), dtype=torch.float64).view(nct, 64)
    med = t.median(0).values
    order = sorted((float(med[k]), k) for k in SLOTS if float(med[k]) > 0)
    res["timeline"] = [f"{int(v):6d} {SLOTS[k]}" for v, k in order]
    res["ctas"] = nct
    return res


def run_tattn(N, B, timed):
    """Temporal attention: tcgen05 kernels (ALPRO_TATTN_TC=1) vs the CUDA-core kernels, T = 8."""
    import torch
    from alpro_b200 import ops
    dev = "cuda"
    T, heads = 8, (12 if timed else 3)
    d = heads * 64
    Sc = 1 + N * T
    gen = torch.Generator(device=dev).manual_seed(11)
    qkv = torch.randn(B * Sc, 3 * d, device=dev, generator=gen).half()
    do = torch.randn(B * Sc, d, device=dev, generator=gen).half()
    res = {}
    outs = {}
    flush = torch.empty(160 * 1024 * 1024, device=dev, dtype=torch.uint8)
    for name, flag in (("cc", "0"), ("tc", "1")):
        os.environ["ALPRO_TATTN_TC"] = flag
        o = torch.full((B * Sc, d), 7.0, device=dev, dtype=torch.float16)
        dqkv = torch.full((B * Sc, 3 * d), 7.0, device=dev, dtype=torch.float16)
        ops.temporal_attn_fwd(qkv, o, B, N, T, heads, 0.125)
        ops.temporal_attn_bwd(qkv, do, dqkv, B, N, T, heads, 0.125)
        torch.cuda.synchronize()
        outs[name] = (o.float(), dqkv.float())
        if timed:
            for tag, fn in (("fwd", lambda: ops.temporal_attn_fwd(qkv, o, B, N, T, heads, 0.125)),
                            ("bwd", lambda: ops.temporal_attn_bwd(qkv, do, dqkv, B, N, T, heads, 0.125))):
                tot = 0.0
                for _ in range(6):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    fn()
                    e1.record()
                    torch.cuda.synchronize()
                    tot += e0.elapsed_time(e1)
                res[f"{tag}_{name}_ms"] = round(tot / 6, 4)
    res["fwd_tc_vs_cc"] = rel(outs["tc"][0], outs["cc"][0])
    for k in ("dq", "dk", "dv"):
        res[f"bwd_tc_vs_cc_{k}"] = rel(blocks(outs["tc"][1], d)[k], blocks(outs["cc"][1], d)[k])
    res["finite"] = bool(torch.isfinite(outs["tc"][0]).all() and torch.isfinite(outs["tc"][1]).all())
    res["cls_zero"] = float(outs["tc"][0].view(B, Sc, d)[:, 0].abs().max() + outs["tc"][1].view(B, Sc, 3 * d)[:, 0].abs().max())
    if timed:
        rows = B * Sc
        res["fwd_tc_GBs"] = round(rows * d * 2 * 4 / (res["fwd_tc_ms"] * 1e-3) / 1e9, 1)
        res["bwd_tc_GBs"] = round(rows * d * 2 * 7 / (res["bwd_tc_ms"] * 1e-3) / 1e9, 1)
    return res


def run_lnbwd(which):
    """Timing of layernorm_bwd at the two large shapes of the step (video token stream / fusion stream with dropout)."""
    import torch
    from alpro_b200 import ops
    dev = "cuda"
    gen = torch.Generator(device=dev).manual_seed(5)
    d = 768
    M = 50208 if which == "video" else 30336
    x = torch.randn(M, d, device=dev, generator=gen)
    mean, rstd = x.mean(1).contiguous(), (1.0 / x.std(1)).contiguous()
    gamma = torch.randn(d, device=dev, generator=gen)
    dx = torch.randn(M, d, device=dev, generator=gen)
    dx16 = torch.empty(M, d, device=dev, dtype=torch.float16)
    dg, db, cs = (torch.zeros(d, device=dev) for _ in range(3))
    if which == "video":
        dy = torch.randn(M, d, device=dev, generator=gen).half()
        rsc = torch.rand(M, device=dev, generator=gen)
        fn = lambda: ops.layernorm_bwd(dy, x, mean, rstd, gamma, dx, 1, dx16=dx16, dgamma=dg, dbeta=db, param_scale=1.0,
                                       colsum=cs, dx16_row_scale=rsc)
        nbytes = M * d * (2 + 4 + 4 + 4 + 2)
    else:
        dy = torch.randn(M, d, device=dev, generator=gen)
        mask = (torch.rand(M, d, device=dev, generator=gen) > 0.1).half()
        fn = lambda: ops.layernorm_bwd(dy, x, mean, rstd, gamma, dx, 0, dx16=dx16, dgamma=dg, dbeta=db, param_scale=1.0,
                                       colsum=cs, dx16_mul16=mask)
        nbytes = M * d * (4 + 4 + 4 + 2 + 2)
    flush = torch.empty(160 * 1024 * 1024, device=dev, dtype=torch.uint8)
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / 6
    return {"M": M, "ms": round(ms, 4), "GBs": round(nbytes / (ms * 1e-3) / 1e9, 1)}


def run_case(case):
    parts = case.split(":")
    if parts[0] == "bert":
        return run_bert(int(parts[1]), parts[2], float(parts[3]))
    if parts[0] == "vit":
        return run_vit(int(parts[1]), int(parts[2]))
    if parts[0] == "time":
        return run_time(parts[1])
    if parts[0] == "trace":
        return run_trace(parts[1])
    if parts[0] == "stagger":   # stagger:<cycles>: time:vit with ALPRO_ATTN_STAGGER set
        os.environ["ALPRO_ATTN_STAGGER"] = parts[1]
        r = run_time("vit")
        return {k: v for k, v in r.items() if k in ("bwd_tc_ms", "bwd_tc_vs_mma")}
    if parts[0] == "lnbwd":
        return run_lnbwd(parts[1])
    if parts[0] == "tattn":
        return run_tattn(int(parts[1]), int(parts[2]), len(parts) > 3)
    if parts[0] == "ftrace":
        return run_ftrace(parts[1])
    raise SystemExit("unknown case " + case)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", help="comma-separated cases to run in this process (one JSON line each)")
    ap.add_argument("--timeout", type=int, default=150, help="per group of cases (seconds)")
    ap.add_argument("--only", default="", help="comma-separated prefixes of the cases to run")
    args = ap.parse_args()
    if args.cases:
        for c in args.cases.split(","):
            try:
                r = run_case(c)
                r = {k: (round(v, 6) if isinstance(v, float) else v) for k, v in r.items()}
                print(json.dumps({"case": c, **r}), flush=True)
            except Exception as e:  # noqa: BLE001  (diagnostic tool: report and carry on)
                print(json.dumps({"case": c, "error": repr(e)[-600:]}), flush=True)
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    logp = os.path.join(ROOT, "gpurun_out", "attn_tc_check.jsonl")
    sel = [c for c in CASES if not args.only or any(c.startswith(p) for p in args.only.split(","))]
    groups = {}
    for c in sel:   # one subprocess per family: a hung kernel only loses the rest of its own group
        groups.setdefault(c.split(":")[0], []).append(c)
    with open(logp, "a") as log:
        for fam, cs in groups.items():
            proc = subprocess.Popen([sys.executable, os.path.abspath(__file__), "--cases", ",".join(cs)],
                                    stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            import threading
            killer = threading.Timer(args.timeout, proc.kill)
            killer.start()
            for line in proc.stdout:
                line = line.rstrip()
                if line.startswith("{"):
                    log.write(line + "\n")
                    log.flush()
                print(line, flush=True)
            rc = proc.wait()
            killer.cancel()
            if rc != 0:
                msg = json.dumps({"group": fam, "error": "exit code %d (timeout / crash)" % rc})
                print(msg, flush=True)
                log.write(msg + "\n")
                sys.exit(2)   # a hang or crash: do not spend GPU time on the remaining groups
    bad = 0
    for line in open(logp):
        try:
            r = json.loads(line)
        except ValueError:
            continue
        if "error" in r or r.get("finite") is False:
            bad += 1
        bad += sum(1 for k, v in r.items() if "_vs_" in k and isinstance(v, float) and not (v < 0.02))
    print("check_attn_tc: %d problem(s)" % bad, flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
