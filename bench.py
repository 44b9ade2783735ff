#!/usr/bin/env python
"""bench.py — video-text pairs/sec for one ALPRO training step (forward + backward incl. the VTC feature exchange and
the gradient all-reduce) on synthetic 8-frame 224^2 clips + 40-token captions (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference ...                     CPU arm: the oracle port of the reference path on host cores

Prints ONE JSON line (rank 0). `value` = whole-job pairs/s with inputs resident in HBM; `e2e` = the same metric through
the reference-facing nn.Module call with HOST (pinned) batches, H2D copies and a D2H read of the losses inside the
timed region. `roofline` describes the dominant kernel (gemm16_kernel, tensor bound); `cpu_baseline` is the oracle
timed on the host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

T_FRAMES, IMG, TXT_LEN, VOCAB, NUM_ENT = 8, 224, 40, 30522, 1000
# algorithmic FLOPs per pair, BASELINE.md §2 (2*M*N*K per GEMM, 4*S*S*d per attention layer; no recompute/padding)
FLOP_PER_PAIR = {"pretrain": 1846.9e9, "retrieval": 1375.7e9}


def full_cfg(kind):
    from alpro_b200 import configs
    bert = dict(configs.BASE_BERT)
    video = dict(configs.BASE_VIDEO)
    video.update(num_frm=T_FRAMES, img_size=IMG)
    vis = dict(d=768, depth=12, heads=12, T=T_FRAMES, img=IMG, patch=16)
    return bert, video, vis


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(bf16_tflops=p.get("bf16_tflops", 1590.0), sustained=p.get("bf16_tflops_sustained", 1400.0),
                    hbm_gbs=p.get("hbm_gbs", 6650.0), source="measured")
    return dict(bf16_tflops=1590.0, sustained=1400.0, hbm_gbs=6650.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.lines = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            parts = [x.strip() for x in l.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_batch(kind, B, seed, device, pinned=False):
    from alpro_b200 import synth
    batch = synth.synth_batch(kind, B, T_FRAMES, IMG, TXT_LEN, VOCAB, seed=seed, num_entities=NUM_ENT)
    out = {}
    for k, v in batch.items():
        if torch.is_tensor(v):
            out[k] = v.pin_memory() if pinned else v.to(device)
        else:
            out[k] = v
    return out


def batch_bytes(batch):
    return sum(v.numel() * v.element_size() for v in batch.values() if torch.is_tensor(v))


def build_model(kind, device, seed=0):
    from alpro_b200 import modeling
    bert, video, vis = full_cfg(kind)
    bert = dict(bert)
    bert["num_entities"] = NUM_ENT
    cls = modeling.AlproForPretrain if kind == "pretrain" else modeling.AlproForVideoTextRetrieval
    torch.manual_seed(seed)
    model = cls(bert, video).to(device)
    g = torch.Generator(device=device).manual_seed(seed)
    with torch.no_grad():   # non-degenerate random weights (reference-style init leaves temporal_fc at zero)
        for n, p in model.named_parameters():
            leaf = n.split(".")[-1]
            parent = n.split(".")[-2].lower() if "." in n else ""
            if n.endswith("temp"):
                p.fill_(0.07)
            elif "norm" in parent:
                p.copy_((1.0 if leaf == "weight" else 0.0) + 0.05 * torch.randn(p.shape, device=device, generator=g))
            else:
                p.copy_(0.02 * torch.randn(p.shape, device=device, generator=g))
    model.train()
    return model


def run_ours(args):
    from alpro_b200 import comm as acomm, ops, _lib
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    kind = args.workload
    B = args.batch
    model = build_model(kind, device)
    if world > 1:
        acomm.attach(model)
    dev_batch = make_batch(kind, B, 1234 + rank, device)
    host_batch = make_batch(kind, B, 1234 + rank, device, pinned=True)
    h2d = batch_bytes(host_batch)

    def step(batch):
        out = model(batch)
        loss = sum(v for k, v in out.items() if k.endswith("_loss") and v is not None)
        loss.backward()
        if world > 1:
            acomm.allreduce_gradients(model)
        for p in model.parameters():
            p.grad = None
        return out

    # End-to-end step through the public API as a trainer drives it: pinned host batches go through
    # alpro_b200.prefetch.PrefetchLoader (the reference's PrefetchLoader, src/datasets/dataloader.py:80-157): the H2D
    # copy of step i+1 runs on a side stream under the kernels of step i. Every timed step issues exactly one batch
    # copy and reads its losses back to the host.
    import itertools
    from alpro_b200.prefetch import PrefetchLoader

    def make_e2e(hb):
        it = iter(PrefetchLoader(itertools.repeat(hb), device))

        def run():
            out = step(next(it))
            vals = torch.stack([out[k].detach() for k in out if k.endswith("_loss") and out[k] is not None])
            return vals.cpu()   # D2H read of the step's result (synchronises)
        return run

    e2e_step = make_e2e(host_batch)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = {}

    def timed(fn, steps, tag=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        if tag:   # host time to ENQUEUE the steps (launch calls are asynchronous): << device time = not launch-bound
            host_ms[tag] = (time.perf_counter() - t0) * 1e3 / steps
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    if args.ncu:   # profiling aid: ncu --profile-from-start off ... python bench.py --ncu   (never a bench value)
        step(dev_batch)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        step(dev_batch)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        return
    for _ in range(max(args.warmup, 3)):
        step(dev_batch)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    calls0 = _lib.counted.calls
    ms_step = timed(lambda: step(dev_batch), args.steps, tag="step")
    launches = (_lib.counted.calls - calls0) // args.steps
    clocks = sampler.stop() if rank == 0 else None
    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)

    # same end-to-end step fed with RAW uint8 frames (ImageNorm fused into the patch gather)
    g8 = torch.Generator().manual_seed(7 + rank)
    host_u8 = dict(host_batch)
    for k in ("visual_inputs", "crop_visual_inputs", "context_visual_inputs"):
        if k in host_u8:
            host_u8[k] = torch.randint(0, 256, tuple(host_batch[k].shape), dtype=torch.uint8, generator=g8).pin_memory()
    h2d_u8 = batch_bytes(host_u8)

    e2e_u8_step = make_e2e(host_u8)

    e2e_u8_step()
    ms_e2e_u8 = timed(e2e_u8_step, args.steps)

    # roofline of the dominant kernel: event-timed GEMM launches of one extra step (tensor bound)
    ops.GEMM_PROFILE = []
    step(dev_batch)
    torch.cuda.synchronize()
    gemm_ms = sum(e0.elapsed_time(e1) for (_, _, _, e0, e1) in ops.GEMM_PROFILE)
    gemm_flop = sum(2.0 * M * N * K for (M, N, K, _, _) in ops.GEMM_PROFILE)
    n_gemm = len(ops.GEMM_PROFILE)
    # algorithmic bytes of the same launches: both 16-bit operands once + 2 bytes per output element (lower bound: fp32
    # outputs / residual reads / the saved GELU derivative add to it)
    gemm_bytes = sum(2.0 * (M * K + N * K + M * N) for (M, N, K, _, _) in ops.GEMM_PROFILE)
    ops.GEMM_PROFILE = None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f)
    peaks = load_peaks()
    achieved = gemm_flop / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0

    # extra (not part of the metric): the same step followed by the fused clip + AdamW update
    ms_opt = None
    if not args.no_optimizer:
        from alpro_b200 import optim
        opt = optim.FusedAdamW(model, lr=2.5e-5, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.0, max_grad_norm=5.0)

        def step_opt():
            out = model(dev_batch)
            sum(v for k, v in out.items() if k.endswith("_loss") and v is not None).backward()
            if world > 1:
                acomm.allreduce_gradients(model)
            opt.step()
            opt.zero_grad()

        step_opt()
        ms_opt = timed(step_opt, max(2, args.steps // 2))

    # sanity of the measured computation: losses and the gradient norm of one more step must be finite
    chk = model(dev_batch)
    chk_loss = sum(v for k, v in chk.items() if k.endswith("_loss") and v is not None)
    chk_loss.backward()
    gflat = model.engine.last_grads.flat
    finite = bool(torch.isfinite(chk_loss).item()) and bool(torch.isfinite(gflat).all().item())
    loss_vals = {k: round(float(v), 5) for k, v in chk.items() if k.endswith("_loss") and v is not None}
    gnorm = float(gflat.double().norm())
    if world > 1:
        acomm.allreduce_gradients(model)
    for p in model.parameters():
        p.grad = None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pairs = B * world
    value = pairs / (ms_step * 1e-3)
    res = {
        "metric": "video-text pairs/sec (fwd+bwd, 8x224^2)", "value": round(value, 3), "unit": "pairs/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp16 operands / fp32 accumulate (tcgen05 kind::f16), fp32 residual+statistics",
        "data": "synthetic", "impl": "alpro_b200",
        "config": {"workload": f"alpro_{kind}_step", "clips_per_gpu": B, "global_batch": pairs, "frames": T_FRAMES,
                   "img": IMG, "txt_len": TXT_LEN, "parallelism": f"dp{world}", "l2": "inputs_exceed_l2",
                   "losses": "VTC+VTM+MLM+PEM" if kind == "pretrain" else "VTC+VTM", "optimizer": "none (fwd+bwd+allreduce)",
                   "mode": "train (BERT dropout 0.1 hidden+attention, DropPath 0.1 active; teacher in eval)"},
        "tensor_frac_of_peak_whole_step": round(FLOP_PER_PAIR[kind] * value / world / (peaks["sustained"] * 1e12), 4),
        "e2e": {"value": round(pairs / (ms_e2e * 1e-3), 3), "unit": "pairs/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 16 if kind == "pretrain" else 8, "ms_per_step": round(ms_e2e, 3),
                "loader": "alpro_b200.prefetch.PrefetchLoader (pinned host batch, side-stream H2D of the next step)"},
        "e2e_uint8_inputs": {"value": round(pairs / (ms_e2e_u8 * 1e-3), 3), "unit": "pairs/s",
                             "h2d_bytes_per_step": h2d_u8, "ms_per_step": round(ms_e2e_u8, 3)},
        "finite": finite, "losses": loss_vals, "grad_norm": round(gnorm, 5),
        "gpu_launches": int(launches),
        "host_enqueue_ms_per_step": round(host_ms.get("step", 0.0), 3),
        "ms_per_step_with_fused_adamw": round(ms_opt, 3) if ms_opt else None,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "gemm16_kernel (tcgen05)", "achieved": round(achieved, 1),
                     "peak": peaks["sustained"], "unit": "TFLOP/s", "frac": round(achieved / peaks["sustained"], 4),
                     "peak_source": peaks["source"] + " (cuBLAS bf16 sustained)", "launches": n_gemm,
                     "gemm_ms_per_step": round(gemm_ms, 3), "gemm_share_of_step": round(gemm_ms / ms_step, 3),
                     "traffic": traffic.get("traffic_bytes_per_launch") if traffic else None,
                     "traffic_unit": "bytes/launch (ncu dram__bytes_read+write, profiles/gemm_traffic.json)",
                     "algorithmic_bytes_per_launch": round(gemm_bytes / max(n_gemm, 1), 1),
                     "flop_per_launch": round(gemm_flop / max(n_gemm, 1), 1)},
    }
    if world == 1 and not args.no_cpu_baseline:
        res["cpu_baseline"] = cpu_baseline(kind, steps=2, B=4, warmup=1)
    print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


def usable_cores(cap=32):
    """Host threads the CPU arm may use: scheduler affinity, clipped by the cgroup CPU quota and by `cap` (torch's
    intra-op scaling on this model is flat beyond ~32 threads; 128 oversubscribed threads ran 30x slower)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return max(1, min(n, cap))


def cpu_baseline(kind, steps=2, B=4, warmup=0):
    """The oracle (CPU restatement of the reference path, pinned to the reference's golden vectors) on the host cores.
    Bounded sample: B clips of the same workload (8x224^2, L=40, full-size model), fwd+bwd."""
    from alpro_b200 import synth
    from oracle import alpro_oracle
    cores = usable_cores()
    torch.set_num_threads(cores)
    bert, video, vis = full_cfg(kind)
    spec = synth.model_spec(kind, bert, vis, NUM_ENT)
    g = torch.Generator().manual_seed(0)
    sd = {}
    for k, shp in spec.items():
        c = synth.canonical_name(k)
        if c in sd:
            sd[k] = sd[c]
        elif k.endswith("position_ids"):
            sd[k] = torch.arange(shp[1]).unsqueeze(0)
        elif k.endswith("temp"):
            sd[k] = torch.tensor(0.07)
        elif k.endswith("prompt_feat"):
            sd[k] = torch.rand(shp, generator=g)
        else:
            sd[k] = 0.02 * torch.randn(shp, generator=g)
            if "norm" in k.lower().split(".")[-2] and k.endswith("weight"):
                sd[k] += 1.0
    for k, v in sd.items():
        if v.is_floating_point() and not k.startswith("prompter."):
            v.requires_grad_(True)
    batch = synth.synth_batch(kind, B, T_FRAMES, IMG, TXT_LEN, VOCAB, seed=99, num_entities=NUM_ENT)
    fwd = alpro_oracle.pretrain_forward if kind == "pretrain" else alpro_oracle.retrieval_forward
    times = []
    for i in range(warmup + steps):
        t0 = time.time()
        out = fwd(sd, bert, vis, batch)
        loss = sum(v for k, v in out.items() if k.endswith("_loss") and v is not None)
        loss.backward()
        if i >= warmup:
            times.append(time.time() - t0)
        for v in sd.values():
            v.grad = None
    t = sum(times) / len(times)
    return {"value": round(B / t, 4), "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": f"{B} pairs per step (8x224^2 clips, L=40, full-size model), fwd+bwd, torch CPU fp32, "
                      f"mean of {steps} steps after {warmup} warm-up",
            "seconds": round(t, 2), "best_seconds": round(min(times), 2)}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; /root/reference does not exist on the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind = args.workload
    # K timed steps after W warm-up steps, each step a bounded sample (4 clips) of the workload: ~5 s per step on the
    # box's host cores, so the default K=5/W=3 run ends within a minute; very large K/W are clamped to stay in minutes
    steps = max(1, min(args.steps, 20))
    warmup = max(0, min(args.warmup, 5))
    Bs = 4
    cb = cpu_baseline(kind, steps=steps, B=Bs, warmup=warmup)
    res = {"impl": "reference", "metric": "video-text pairs/sec (fwd+bwd, 8x224^2)", "value": cb["value"],
           "unit": "pairs/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": steps, "warmup": warmup,
           "ms_per_step": round(1e3 * cb["seconds"], 1), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"alpro_{kind}_step", "clips_per_step": Bs, "frames": T_FRAMES, "img": IMG,
                      "txt_len": TXT_LEN, "losses": "VTC+VTM+MLM+PEM" if kind == "pretrain" else "VTC+VTM",
                      "optimizer": "none (fwd+bwd)", "mode": "eval-mode math (no dropout / DropPath draws)",
                      "note": "bounded sample of the same workload on host cores (the reference's torch CPU path as "
                              "restated by oracle/alpro_oracle.py)"},
           "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(res), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="pretrain", choices=["pretrain", "retrieval"])
    ap.add_argument("--batch", type=int, default=32, help="clips per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-optimizer", action="store_true", help="skip the extra fwd+bwd+FusedAdamW timing")
    ap.add_argument("--ncu", action="store_true", help="one warm step, then one step between cudaProfilerStart/Stop")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun when invoked directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", __file__] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
