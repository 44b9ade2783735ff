"""GPU probe for alpro_gemm16: runs each case in its own subprocess (a hang or fault in one case cannot take the
others down) and writes gpurun_out/probe_gemm.json. Usage on the GPU box: python tools/probe_gemm.py"""
import json
import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    # name: (M, N, K, a_layout, b_layout, a_dt, b_dt, extras)
    "nt_1tile": (128, 256, 64, 0, 0, "f16", "f16", {}),
    "nt_k768": (128, 256, 768, 0, 0, "f16", "f16", {}),
    "nt_multi": (1024, 768, 768, 0, 0, "f16", "f16", {}),
    "nt_tails": (200, 296, 104, 0, 0, "f16", "f16", {}),
    "nt_bf16": (512, 512, 256, 0, 0, "bf16", "bf16", {}),
        "nt_bias_gelu": (640, 3072, 768, 0, 0, "f16", "f16", {"bias": 1, "act": 1, "out16b": 1}),
    "nt_resid_skip": (3 * 17, 768, 768, 0, 0, "f16", "f16", {"bias": 1, "resid": 1, "skip": 17, "out32": 1}),
    "nn_dgrad": (1024, 768, 2304, 0, 1, "bf16", "bf16", {}),
    "nn_dgrad_gelugrad": (640, 768, 3072, 0, 1, "f16", "f16", {"act": 2}),
    "nn_tails": (200, 296, 104, 0, 1, "f16", "f16", {}),
    "tn_wgrad": (2304, 768, 4096, 1, 1, "bf16", "bf16", {"split": -1}),
    "tn_wgrad_f16": (768, 3072, 5000, 1, 1, "f16", "f16", {"split": -1}),
    "tn_nosplit": (256, 512, 1000, 1, 1, "f16", "f16", {"out32": 1}),
    "tk_amn_bk": (256, 512, 1000, 1, 0, "f16", "f16", {"out32": 1}),
    "nt_vocab_unaligned": (1280, 30522, 768, 0, 0, "f16", "f16", {"bias": 1, "out32": 1}),
    "perf_qkv": (50176, 2304, 768, 0, 0, "f16", "f16", {"bias": 1, "perf": 1}),
    "perf_fc2": (50208, 768, 3072, 0, 0, "f16", "f16", {"bias": 1, "resid": 1, "out32": 1, "perf": 1}),
    "perf_proj_resid": (50208, 768, 768, 0, 0, "f16", "f16", {"bias": 1, "resid": 1, "out32": 1, "perf": 1}),
    "nt_resid_ragged": (1000, 800, 328, 0, 0, "f16", "f16", {"bias": 1, "resid": 1, "skip": 7, "out32": 1}),
    "nt_gelu_nosave": (777, 1024, 256, 0, 0, "bf16", "bf16", {"bias": 1, "act": 1}),
    "perf_wgrad_sq": (768, 768, 50208, 1, 1, "f16", "f16", {"split": -1, "perf": 1}),
    "perf_wgrad_fc2": (768, 3072, 50208, 1, 1, "f16", "f16", {"split": -1, "perf": 1}),
    "perf_wgrad": (2304, 768, 50176, 1, 1, "f16", "f16", {"split": -1, "perf": 1}),
    "perf_fc1_gelu": (50208, 3072, 768, 0, 0, "f16", "f16", {"bias": 1, "act": 1, "out16b": 1, "perf": 1}),
    "perf_dgrad": (50176, 768, 2304, 0, 1, "f16", "f16", {"perf": 1}),
    "perf_dgrad_gelugrad": (50208, 3072, 768, 0, 1, "f16", "f16", {"act": 2, "perf": 1}),
    "perf_dgrad_sq": (50208, 768, 768, 0, 1, "f16", "f16", {"perf": 1}),          # proj / temporal_fc dgrad (B MN-major)
    "perf_sq_kmajor": (50208, 768, 768, 0, 0, "f16", "f16", {"perf": 1}),         # same shape, B K-major
    "perf_sq_n1536": (50208, 1536, 768, 0, 1, "f16", "f16", {"perf": 1}),
}


def run_case(name):
    import torch
    from alpro_b200 import ops
    M, N, K, al, bl, adt, bdt, ex = CASES[name]
    dt = {"f16": torch.float16, "bf16": torch.bfloat16}
    g = torch.Generator(device="cuda").manual_seed(1234)
    dev = "cuda"
    a_shape = (M, K) if al == 0 else (K, M)
    b_shape = (N, K) if bl == 0 else (K, N)

    def mk(shape, dtype):
        ld = (shape[1] + 7) // 8 * 8
        buf = torch.zeros(shape[0], ld, device=dev, dtype=dtype)
        buf[:, :shape[1]] = (torch.randn(shape, device=dev, generator=g) * 0.5).to(dtype)
        return buf[:, :shape[1]]

    a = mk(a_shape, dt[adt])
    b = mk(b_shape, dt[bdt])
    A = a.float() if al == 0 else a.float().t()
    B = b.float() if bl == 0 else b.float().t()
    ref = A @ B.t()  # [M,N]
    kw = {}
    bias = None
    if ex.get("bias"):
        bias = torch.randn(N, device=dev, generator=g)
        kw["bias"] = bias
        ref = ref + bias
    act = ex.get("act", 0)
    aux = None
    pre = ref.clone()
    if act == 1:
        ref = torch.nn.functional.gelu(ref)
    elif act == 2:
        aux = (torch.randn(M, N, device=dev, generator=g)).to(torch.float16)
        ref = ref * aux.float()
        kw["aux"] = aux
    kw["act"] = act
    if ex.get("resid"):
        resid = torch.randn(M, N, device=dev, generator=g)
        kw["resid"] = resid
        sk = ex.get("skip", 0)
        out = ref + resid
        if sk:
            rows = torch.arange(M, device=dev) % sk == 0
            out[rows] = resid[rows]
            kw["skip_period"] = sk
        ref = out
    split = ex.get("split", 0)
    use32 = ex.get("out32") or split != 0
    out16b = None
    if use32:
        if N % 4 == 0:
            out = torch.zeros(M, N, device=dev)
        else:
            out = torch.zeros(M * N + 4, device=dev)[1:1 + M * N].view(M, N)  # deliberately unaligned
        kw["out32"] = out
    else:
        ldn = (N + 7) // 8 * 8
        out = torch.zeros(M, ldn, device=dev, dtype=torch.float16)[:, :N]
        kw["out16"] = out
    if ex.get("out16b"):
        out16b = torch.zeros(M, N, device=dev, dtype=torch.float16)
        kw["out16b"] = out16b
    kw["split_k"] = split
    ops.gemm16(a, b, a_layout=al, b_layout=bl, **kw)
    torch.cuda.synchronize()
    res = {"case": name, "M": M, "N": N, "K": K}
    o = out.float()
    err = (o - ref).abs().max().item()
    scale = ref.abs().max().item()
    res["max_abs_err"] = err
    res["ref_max"] = scale
    res["rel"] = err / max(scale, 1e-9)
    tol = 2e-3 if not use32 else 2e-5 * math.sqrt(K)
    res["ok"] = bool(res["rel"] < tol)
    if out16b is not None:
        u = pre.clone().requires_grad_(True)
        torch.nn.functional.gelu(u).sum().backward()
        e2 = (out16b.float() - u.grad).abs().max().item() / max(u.grad.abs().max().item(), 1e-9)
        res["pre_rel"] = e2
        res["ok"] = res["ok"] and e2 < 2e-3
    if ex.get("perf"):
        if split:
            pass
        for _ in range(3):
            ops.gemm16(a, b, a_layout=al, b_layout=bl, **kw)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        iters = 10
        e0.record()
        for _ in range(iters):
            ops.gemm16(a, b, a_layout=al, b_layout=bl, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        res["ms"] = ms
        res["tflops"] = 2.0 * M * N * K / ms / 1e9
        # cuBLAS comparison for context
        A16 = A.to(torch.bfloat16).contiguous()
        B16 = B.to(torch.bfloat16).contiguous()
        for _ in range(3):
            A16 @ B16.t()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            A16 @ B16.t()
        e1.record()
        torch.cuda.synchronize()
        res["cublas_tflops"] = 2.0 * M * N * K / (e0.elapsed_time(e1) / iters) / 1e9
    print("RESULT " + json.dumps(res))


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--case":
        run_case(sys.argv[2])
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    names = sys.argv[1:] or list(CASES)
    results = []
    for name in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, "--case", name], capture_output=True, text=True, timeout=180)
            line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            if line:
                res = json.loads(line[-1][7:])
            else:
                res = {"case": name, "ok": False, "rc": r.returncode, "stderr": r.stderr[-1500:]}
        except subprocess.TimeoutExpired:
            res = {"case": name, "ok": False, "timeout": True}
        res["wall_s"] = round(time.time() - t0, 1)
        print(json.dumps(res), flush=True)
        results.append(res)
        with open(os.path.join(ROOT, "gpurun_out", "probe_gemm.json"), "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
