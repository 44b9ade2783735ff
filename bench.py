#!/usr/bin/env python
"""bench.py — video-text pairs/sec for one ALPRO training step (forward + backward incl. the VTC feature exchange and
the gradient all-reduce) on synthetic 8-frame 224^2 clips + 40-token captions (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference ...                     CPU arm: the oracle port of the reference path on host cores
  python bench.py --workload retrieval|timesformer_fwd|vtc_allgather     the other BASELINE.json configs

Prints ONE JSON line (rank 0). `value` = whole-job pairs/s with inputs resident in HBM; `e2e` = the same metric through
the reference-facing nn.Module call with HOST (pinned) batches, H2D copies and a D2H read of the losses inside the
timed region. `roofline` describes the dominant kernel (gemm16, tensor bound); `cpu_baseline` is the oracle timed on the
host cores on a bounded sample; `torch_gpu_baseline` is the reference's PyTorch path (oracle restatement, eager fp32
with TF32 off = the reference's default, and bf16 autocast) timed on the SAME B200 after the timed region;
`dp_parity` (N > 1) is the data-parallel step of the tiny parity config checked against the CPU oracle of the global
objective, outside the timed region.
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
# algorithmic FLOPs per pair, BASELINE.md §2 / SURVEY.md §8d (2*M*N*K per GEMM, 4*S*S*d per attention layer; no
# recompute/padding); per clip for the TimeSformer forward alone
FLOP_PER_PAIR = {"pretrain": 1846.9e9, "retrieval": 1375.7e9}
FLOP_PER_CLIP_FWD = 391.7e9
METRIC = "video-text pairs/sec (fwd+bwd, 8x224^2)"


def full_cfg(kind=None):
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
        sm, mx, pw, reasons = [], [], [], set()
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
            try:
                pw.append(float(parts[2]))
            except ValueError:
                pass
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        pw.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_median": pw[len(pw) // 2] if pw else None}


class EnergyMeter:
    """Board energy over a timed region from the NVML energy counter (mJ since driver load): joules per step and the
    average power against the enforced power limit. avg_w ~ limit_w with `sw_power_cap` active means the step is
    energy-bound: its duration is joules / limit, whatever the kernel durations add up to."""

    def __init__(self, index=0):
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.limit_w = pynvml.nvmlDeviceGetEnforcedPowerLimit(self.h) / 1e3
        except Exception:
            self.h = None

    def read(self):
        if self.h is None:
            return None
        try:
            return self.nv.nvmlDeviceGetTotalEnergyConsumption(self.h) / 1e3, time.perf_counter()
        except Exception:
            return None

    def report(self, a, b, steps, flop_per_step=None):
        if a is None or b is None or b[1] <= a[1]:
            return None
        joules, secs = b[0] - a[0], b[1] - a[1]
        out = {"joules_per_step": round(joules / steps, 2), "avg_w": round(joules / secs, 1),
               "limit_w": round(self.limit_w, 1), "window_s": round(secs, 3), "steps": steps,
               "ms_per_step_in_window": round(secs / steps * 1e3, 3),
               "source": "nvmlDeviceGetTotalEnergyConsumption around a synchronised run of the same step"}
        if flop_per_step:
            out["pj_per_flop_whole_step"] = round(joules / steps / flop_per_step * 1e12, 3)
        return out


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


def build_model(kind, device, seed=0, dtype=None):
    from alpro_b200 import modeling
    bert, video, vis = full_cfg(kind)
    bert = dict(bert)
    bert["num_entities"] = NUM_ENT
    cls = modeling.AlproForPretrain if kind == "pretrain" else modeling.AlproForVideoTextRetrieval
    torch.manual_seed(seed)
    model = cls(bert, video).to(device)
    if dtype is not None:
        model.set_compute_dtype(dtype)
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


def _dtype_arg(name):
    return {"fp16": torch.float16, "bf16": torch.bfloat16}[name]


DTYPE_STR = {torch.float16: "fp16 operands / fp32 accumulate (tcgen05 kind::f16), fp32 residual+statistics",
             torch.bfloat16: "bf16 operands / fp32 accumulate (tcgen05 kind::f16), fp32 residual+statistics"}


class Dist:
    """Process-group plumbing of one bench run (one rank per GPU under torchrun)."""

    def __init__(self):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.device = torch.device("cuda", self.local)
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.device)

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps, host_ms=None, tag=None):
        """K calls of fn bracketed by barrier + synchronize, CUDA events on the launching stream, MAX over ranks."""
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        if tag and host_ms is not None:
            host_ms[tag] = (time.perf_counter() - t0) * 1e3 / steps
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.device)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    def close(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


# =====================================================================================================================
# training-step workloads (BASELINE configs[2] pretrain, configs[3] retrieval)
# =====================================================================================================================
def run_ours(args):
    from alpro_b200 import comm as acomm, ops, _lib
    D = Dist()
    world, rank, device = D.world, D.rank, D.device
    kind = args.workload
    B = args.batch
    dtype = _dtype_arg(args.dtype)
    model = build_model(kind, device, dtype=dtype)
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
    host_ms = {}

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
    sampler = ClockSampler(D.local)
    if rank == 0:
        sampler.start()
    calls0 = _lib.counted.calls
    ms_step = D.timed(lambda: step(dev_batch), args.steps, host_ms, "step")
    launches = (_lib.counted.calls - calls0) // args.steps
    clocks = sampler.stop() if rank == 0 else None
    # board energy of the same step, measured after the timed region over a longer window (the NVML counter updates every
    # ~100 ms: 30+ steps keep the two end-point errors below a few percent); the window is bracketed by synchronize()
    # (A step contains the data-parallel collectives: anything that runs extra steps must run them on EVERY rank.)
    # Single-GPU runs only: the scaling runs keep exactly the control flow that was validated at 8 GPUs.
    power = None
    if world == 1:
        meter = EnergyMeter(D.local)
        n_en = max(args.steps, 30)
        D.barrier()
        en0 = meter.read()
        for _ in range(n_en):
            step(dev_batch)
        D.barrier()
        en1 = meter.read()
        power = meter.report(en0, en1, n_en, FLOP_PER_PAIR[kind] * B)
    e2e_step()
    ms_e2e = D.timed(e2e_step, args.steps)

    # same end-to-end step fed with RAW uint8 frames (ImageNorm fused into the patch gather)
    g8 = torch.Generator().manual_seed(7 + rank)
    host_u8 = dict(host_batch)
    for k in ("visual_inputs", "crop_visual_inputs", "context_visual_inputs"):
        if k in host_u8:
            host_u8[k] = torch.randint(0, 256, tuple(host_batch[k].shape), dtype=torch.uint8, generator=g8).pin_memory()
    h2d_u8 = batch_bytes(host_u8)
    e2e_u8_step = make_e2e(host_u8)
    e2e_u8_step()
    ms_e2e_u8 = D.timed(e2e_u8_step, args.steps)

    # Host-side cost of a step, measured instead of inferred: the SAME step at 1 clip per GPU issues the same launches
    # (identical Python / ctypes / descriptor work) with ~1/32 of the device work, so its step time is an upper bound of
    # the per-step host cost. ms_per_step well above it = the device, not the host, sets the pace.
    one_batch = make_batch(kind, 1, 4321 + rank, device)
    for _ in range(2):
        step(one_batch)
    ms_b1 = D.timed(lambda: step(one_batch), max(3, args.steps), host_ms, "b1")

    # roofline of the dominant kernel: event-timed GEMM launches of one extra step (tensor bound)
    ops.GEMM_PROFILE = []
    step(dev_batch)
    torch.cuda.synchronize()
    prof = ops.GEMM_PROFILE
    ops.GEMM_PROFILE = None
    gemm_ms = sum(r[3].elapsed_time(r[4]) for r in prof)
    gemm_flop = sum(2.0 * r[0] * r[1] * r[2] for r in prof)
    gemm_bytes = sum(r[5] for r in prof)
    n_gemm = len(prof)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f)
    peaks = load_peaks()
    achieved = gemm_flop / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0

    # extra (not part of the metric): the same step followed by the fused clip + AdamW update
    ms_opt = None
    opt = None
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
        ms_opt = D.timed(step_opt, max(2, args.steps // 2))

    # sanity of the measured computation: losses and the gradient norm of one more step must be finite
    chk = model(dev_batch)
    chk_loss = sum(v for k, v in chk.items() if k.endswith("_loss") and v is not None)
    chk_loss.backward()
    gflat = model.engine.last_grads.flat
    finite = bool(torch.isfinite(chk_loss).item()) and bool(torch.isfinite(gflat).all().item())
    loss_vals = {k: round(float(v.detach()), 5) for k, v in chk.items() if k.endswith("_loss") and v is not None}
    gnorm = float(gflat.double().norm())
    if world > 1:
        acomm.allreduce_gradients(model)
    for p in model.parameters():
        p.grad = None

    # data-parallel parity, visible to the driver: tiny config, every rank, checked against the CPU oracle
    dp = None
    if world > 1 and not args.no_dp_parity:
        from tests import dp_parity
        dp = dp_parity.run(device, world, rank)
    if rank != 0:
        D.close()
        return
    pairs = B * world
    value = pairs / (ms_step * 1e-3)
    res = {
        "metric": METRIC, "value": round(value, 3), "unit": "pairs/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": DTYPE_STR[dtype], "data": "synthetic", "impl": "alpro_b200",
        "config": step_config(kind, B, pairs, world),
        "tensor_frac_of_peak_whole_step": round(FLOP_PER_PAIR[kind] * value / world / (peaks["sustained"] * 1e12), 4),
        "e2e": {"value": round(pairs / (ms_e2e * 1e-3), 3), "unit": "pairs/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 16 if kind == "pretrain" else 8, "ms_per_step": round(ms_e2e, 3),
                "loader": "alpro_b200.prefetch.PrefetchLoader (pinned host batch, side-stream H2D of the next step)"},
        "e2e_uint8_inputs": {"value": round(pairs / (ms_e2e_u8 * 1e-3), 3), "unit": "pairs/s",
                             "h2d_bytes_per_step": h2d_u8, "ms_per_step": round(ms_e2e_u8, 3)},
        "finite": finite, "losses": loss_vals, "grad_norm": round(gnorm, 5),
        "gpu_launches": int(launches),
        "host": {"enqueue_ms_per_step": round(host_ms.get("step", 0.0), 3),
                 "ms_per_step_at_1_clip": round(ms_b1, 3),
                 "note": "enqueue time includes blocking on a full launch queue; the 1-clip step issues the same "
                         "launches with ~1/32 of the device work = upper bound of the per-step host cost"},
        "ms_per_step_with_fused_adamw": round(ms_opt, 3) if ms_opt else None,
        "clocks": clocks,
        "power": power,
        "roofline": {"bound": "tensor", "kernel": "gemm16 (tcgen05, all launches of one step)",
                     "achieved": round(achieved, 1),
                     "peak": peaks["sustained"], "unit": "TFLOP/s", "frac": round(achieved / peaks["sustained"], 4),
                     "peak_source": peaks["source"] + " (cuBLAS bf16 sustained)", "launches": n_gemm,
                     "gemm_ms_per_step": round(gemm_ms, 3), "gemm_share_of_step": round(gemm_ms / ms_step, 3),
                     "traffic": traffic.get("traffic_bytes_per_launch") if traffic else None,
                     "traffic_source": (f"profiles/gemm_traffic.json tag {traffic.get('tag')} "
                                        f"({traffic.get('launches')} launches)") if traffic else None,
                     "traffic_unit": "bytes/launch (ncu dram__bytes_read+write)",
                     "algorithmic_bytes_per_launch": round(gemm_bytes / max(n_gemm, 1), 1),
                     "algorithmic_bytes_note": "16-bit operands once + every epilogue stream at its real width (fp32 "
                                               "outputs and residual 4 B, 16-bit outputs / aux 2 B)",
                     "flop_per_launch": round(gemm_flop / max(n_gemm, 1), 1)},
    }
    if dp is not None:
        res["dp_parity"] = dp
    # free our arm before the baselines (the torch path needs the HBM)
    del model, dev_batch, host_batch, host_u8, one_batch, chk, gflat, opt, e2e_step, e2e_u8_step
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    if world == 1 and not args.no_torch_baseline:
        try:
            res["torch_gpu_baseline"] = torch_gpu_baseline(kind, device, B, value)
        except Exception as e:          # a baseline must never cost the bench line
            res["torch_gpu_baseline"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if world == 1 and not args.no_cpu_baseline:
        res["cpu_baseline"] = cpu_baseline(kind, steps=2, B=4, warmup=1)
    print(json.dumps(res), flush=True)
    D.close()


def step_config(kind, B, pairs, world):
    return {"workload": f"alpro_{kind}_step", "clips_per_gpu": B, "global_batch": pairs, "frames": T_FRAMES,
            "img": IMG, "txt_len": TXT_LEN, "parallelism": f"dp{world}", "l2": "inputs_exceed_l2",
            "losses": "VTC+VTM+MLM+PEM" if kind == "pretrain" else "VTC+VTM", "optimizer": "none (fwd+bwd+allreduce)",
            "mode": "train (BERT dropout 0.1 hidden+attention, DropPath 0.1 active; teacher in eval)"}


# =====================================================================================================================
# baselines: the reference's PyTorch path (oracle restatement) on host cores and on the same GPU
# =====================================================================================================================
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


def oracle_state_dict(kind, device="cpu", seed=0):
    """Random weights of the full-size model in the reference's state_dict schema, as autograd leaves."""
    from alpro_b200 import synth
    bert, video, vis = full_cfg(kind)
    spec = synth.model_spec(kind, bert, vis, NUM_ENT)
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in spec.items():
        c = synth.canonical_name(k)
        if c in sd:
            sd[k] = sd[c]
        elif k.endswith("position_ids"):
            sd[k] = torch.arange(shp[1]).unsqueeze(0).to(device)
        elif k.endswith("temp"):
            sd[k] = torch.tensor(0.07, device=device)
        elif k.endswith("prompt_feat"):
            sd[k] = torch.rand(shp, generator=g).to(device)
        else:
            v = 0.02 * torch.randn(shp, generator=g)
            if "norm" in k.lower().split(".")[-2] and k.endswith("weight"):
                v += 1.0
            sd[k] = v.to(device)
    for k, v in sd.items():
        if v.is_floating_point() and not k.startswith("prompter."):
            v.requires_grad_(True)
    return sd, bert, vis


def oracle_step(kind, sd, bert, vis, batch, sampler=None):
    """One fwd+bwd of the reference path in TRAIN mode (torch-drawn dropout / DropPath, as model.train() does)."""
    from oracle import alpro_oracle
    fwd = alpro_oracle.pretrain_forward if kind == "pretrain" else alpro_oracle.retrieval_forward
    kw = dict(train=alpro_oracle.random_train(bert, vis, drop_path_rate=0.1))
    if sampler is not None:
        kw["sampler"] = sampler
    out = fwd(sd, bert, vis, batch, **kw)
    loss = sum(v for k, v in out.items() if k.endswith("_loss") and v is not None)
    loss.backward()
    for v in sd.values():
        v.grad = None
    return out


def cpu_baseline(kind, steps=2, B=4, warmup=0):
    """The oracle (CPU restatement of the reference path, pinned to the reference's golden vectors) on the host cores.
    Bounded sample: B clips of the same workload (8x224^2, L=40, full-size model), fwd+bwd in train mode."""
    from alpro_b200 import synth
    cores = usable_cores()
    torch.set_num_threads(cores)
    sd, bert, vis = oracle_state_dict(kind)
    batch = synth.synth_batch(kind, B, T_FRAMES, IMG, TXT_LEN, VOCAB, seed=99, num_entities=NUM_ENT)
    times = []
    for i in range(warmup + steps):
        t0 = time.time()
        oracle_step(kind, sd, bert, vis, batch)
        if i >= warmup:
            times.append(time.time() - t0)
    t = sum(times) / len(times)
    return {"value": round(B / t, 4), "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": f"{B} pairs per step (8x224^2 clips, L=40, full-size model), fwd+bwd in train mode, torch CPU "
                      f"fp32, mean of {steps} steps after {warmup} warm-up",
            "seconds": round(t, 2), "best_seconds": round(min(times), 2)}


def torch_gpu_baseline(kind, device, B, our_value):
    """The reference's PyTorch path on THIS GPU (north_star: ">= 6x per-GPU pairs/sec over the reference PyTorch path"):
    the oracle's modules-as-functions moved to cuda — cuBLAS / cuDNN / eager elementwise kernels, autograd, per-row
    torch.multinomial(...).item() hard-negative draws exactly as alpro_models.py:301-316 — fwd+bwd in train mode on the
    same batch shape. Two settings: eager fp32 with TF32 matmuls off (PyTorch's and therefore the reference's default,
    SURVEY App. B) and bf16 autocast. Runs after our timed region with our arm freed."""
    import gc
    from alpro_b200 import synth
    res = {"impl": "oracle/alpro_oracle.py restatement of src/modeling/alpro_models.py:79-183 on cuda (torch "
                   f"{torch.__version__}, cuBLAS/cuDNN)", "mode": "train", "unit": "pairs/s"}
    sampler = lambda w: int(torch.multinomial(w, 1).item())
    torch.backends.cuda.matmul.allow_tf32 = False          # torch default; stated for the record
    torch.backends.cudnn.allow_tf32 = True
    for name, ctx in (("fp32_eager_tf32_off", None), ("bf16_autocast", torch.bfloat16)):
        b = B
        res[name] = {"error": "out of memory at 1 clip"}
        while b >= 1:
            sd = batch = None
            try:
                sd, bert, vis = oracle_state_dict(kind, device)
                batch = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in
                         synth.synth_batch(kind, b, T_FRAMES, IMG, TXT_LEN, VOCAB, seed=99, num_entities=NUM_ENT).items()}

                def one():
                    if ctx is None:
                        oracle_step(kind, sd, bert, vis, batch, sampler)
                    else:
                        with torch.autocast("cuda", dtype=ctx):
                            oracle_step(kind, sd, bert, vis, batch, sampler)
                one()
                torch.cuda.synchronize()
                n = 3
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(n):
                    one()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / n
                res[name] = {"value": round(b / (ms * 1e-3), 3), "ms_per_step": round(ms, 2), "clips_per_step": b,
                             "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 1e9, 1),
                             "ours_over_this": round(our_value / (b / (ms * 1e-3)), 2)}
                break
            except torch.OutOfMemoryError:
                b //= 2
            finally:
                sd = batch = one = None
                gc.collect()
                torch.cuda.empty_cache()
                torch.cuda.reset_peak_memory_stats()
    return res


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; /root/reference does not exist on the GPU box) on the
    box's host cores, on OUR arm's config / metric / unit, each step a bounded sample of that workload."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    kind = args.workload if args.workload in FLOP_PER_PAIR else "pretrain"
    # K timed steps after W warm-up steps, each step a bounded sample (4 clips) of the workload: ~5 s per step on the
    # box's host cores, so the default K=5/W=3 run ends within a minute; very large K/W are clamped to stay in minutes
    steps = max(1, min(args.steps, 20))
    warmup = max(0, min(args.warmup, 5))
    Bs = 4
    cb = cpu_baseline(kind, steps=steps, B=Bs, warmup=warmup)
    res = {"impl": "reference", "metric": METRIC, "value": cb["value"],
           "unit": "pairs/s", "n_gpus": world, "steps": steps, "warmup": warmup,
           "ms_per_step": round(1e3 * cb["seconds"], 1), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": step_config(kind, args.batch, args.batch * world, world),
           "sample_note": f"each step is a bounded sample of the config's workload: {Bs} of its {args.batch} clips per "
                          "GPU through the same full-size model in train mode on the host cores (oracle/alpro_oracle.py, "
                          "the reference's torch CPU path restated and pinned to its golden vectors)",
           "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(res), flush=True)


# =====================================================================================================================
# BASELINE configs[1]: TimeSformer-B/16 divided space-time forward only, 8x224^2, batch 16, one B200
# =====================================================================================================================
def run_timesformer_fwd(args):
    from alpro_b200 import ops, _lib
    D = Dist()
    device, rank, world = D.device, D.rank, D.world
    B = args.batch if args.batch != 32 else 16
    peaks = load_peaks()
    g = torch.Generator().manual_seed(1234 + rank)
    host = torch.randn(B, T_FRAMES, 3, IMG, IMG, generator=g).pin_memory()
    frames = host.to(device)
    out_host = torch.empty(B, 197, 768).pin_memory()
    results = {}
    for dname in ("bf16", "fp16"):
        model = build_model("retrieval", device, dtype=_dtype_arg(dname)).eval()
        P = model._tensor_dict()

        def fwd():
            with torch.no_grad():
                return model.engine.visual_features(P, frames)

        def e2e():
            with torch.no_grad():
                dev = host.to(device, non_blocking=True)
                out_host.copy_(model._forward_visual_embeds(dev), non_blocking=True)
                torch.cuda.synchronize()
        for _ in range(max(args.warmup, 3)):
            fwd()
        sampler = ClockSampler(D.local)
        if rank == 0 and dname == args.dtype:
            sampler.start()
        calls0 = _lib.counted.calls
        ms = D.timed(fwd, args.steps)
        launches = (_lib.counted.calls - calls0) // args.steps
        clocks = sampler.stop() if rank == 0 and dname == args.dtype else None
        e2e()
        ms_e2e = D.timed(e2e, args.steps)
        ops.GEMM_PROFILE = []
        fwd()
        torch.cuda.synchronize()
        prof, ops.GEMM_PROFILE = ops.GEMM_PROFILE, None
        gemm_ms = sum(r[3].elapsed_time(r[4]) for r in prof)
        gemm_flop = sum(2.0 * r[0] * r[1] * r[2] for r in prof)
        results[dname] = dict(ms=ms, ms_e2e=ms_e2e, launches=launches, clocks=clocks, gemm_ms=gemm_ms,
                              gemm_flop=gemm_flop, n_gemm=len(prof), gemm_bytes=sum(r[5] for r in prof))
        del model, P
        torch.cuda.empty_cache()
    if rank != 0:
        D.close()
        return
    r = results[args.dtype]
    other = results["fp16" if args.dtype == "bf16" else "bf16"]
    clips = B * world
    value = clips / (r["ms"] * 1e-3)
    whole = FLOP_PER_CLIP_FWD * B / (r["ms"] * 1e-3) / 1e12
    res = {"metric": "video clips/sec (TimeSformer-B/16 divided space-time forward, 8x224^2)", "value": round(value, 2),
           "unit": "clips/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
           "ms_per_step": round(r["ms"], 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": DTYPE_STR[_dtype_arg(args.dtype)], "data": "synthetic", "impl": "alpro_b200",
           "config": {"workload": "timesformer_b16_divst_fwd", "clips_per_gpu": B, "frames": T_FRAMES, "img": IMG,
                      "l2": "inputs_exceed_l2", "mode": "eval, no_grad (forward_features + temporal pooling, vit.py:475-503)"},
           "e2e": {"value": round(clips / (r["ms_e2e"] * 1e-3), 2), "unit": "clips/s",
                   "h2d_bytes_per_step": host.numel() * 4, "d2h_bytes_per_step": out_host.numel() * 4,
                   "ms_per_step": round(r["ms_e2e"], 3)},
           "gpu_launches": int(r["launches"]), "clocks": r["clocks"],
           "roofline": {"bound": "tensor", "kernel": "whole divided-ST forward (all kernels, algorithmic FLOPs 391.7 G/clip)",
                        "achieved": round(whole, 1), "peak": peaks["sustained"], "unit": "TFLOP/s",
                        "frac": round(whole / peaks["sustained"], 4),
                        "frac_of_nominal_2250": round(whole / 2250.0, 4),
                        "peak_source": peaks["source"] + " (cuBLAS bf16 sustained)",
                        "gemm_only": {"achieved": round(r["gemm_flop"] / (r["gemm_ms"] * 1e-3) / 1e12, 1),
                                      "launches": r["n_gemm"], "gemm_ms": round(r["gemm_ms"], 3),
                                      "share_of_step": round(r["gemm_ms"] / r["ms"], 3),
                                      "algorithmic_bytes_per_launch": round(r["gemm_bytes"] / max(r["n_gemm"], 1), 1)},
                        "traffic": None},
           "other_dtype": {"dtype": "fp16" if args.dtype == "bf16" else "bf16", "ms_per_step": round(other["ms"], 3),
                           "value": round(clips / (other["ms"] * 1e-3), 2),
                           "tflops": round(FLOP_PER_CLIP_FWD * B / (other["ms"] * 1e-3) / 1e12, 1)}}
    if world == 1 and not args.no_torch_baseline:
        try:
            res["torch_gpu_baseline"] = torch_fwd_baseline(device, B, value)
        except Exception as e:
            res["torch_gpu_baseline"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    print(json.dumps(res), flush=True)
    D.close()


def torch_fwd_baseline(device, B, our_value):
    """The reference's TimeSformer forward (oracle.visual_forward = vit.py:321-377, 475-503) on this GPU, no_grad."""
    from oracle import alpro_oracle
    sd, bert, vis = oracle_state_dict("retrieval", device)
    frames = torch.randn(B, T_FRAMES, 3, IMG, IMG, device=device)
    res = {"unit": "clips/s"}
    for name, ctx in (("fp32_eager_tf32_off", None), ("bf16_autocast", torch.bfloat16)):
        def one():
            with torch.no_grad():
                if ctx is None:
                    return alpro_oracle.visual_forward(sd, "visual_encoder.model.", frames, vis)
                with torch.autocast("cuda", dtype=ctx):
                    return alpro_oracle.visual_forward(sd, "visual_encoder.model.", frames, vis)
        one()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            one()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        res[name] = {"value": round(B / (ms * 1e-3), 2), "ms_per_step": round(ms, 2),
                     "ours_over_this": round(our_value / (B / (ms * 1e-3)), 2)}
    return res


# =====================================================================================================================
# BASELINE configs[4]: VTC contrastive feature exchange, global batch 256/512/1024 at 2/4/8 GPUs
# =====================================================================================================================
def run_vtc_allgather(args):
    """The exchange step of the VTC head in isolation (alpro_models.py:110-111, 764-768): all-gather of the normalised
    [B,256] video | text features as ONE [B,512] message and the backward reduce-scatter of d(gathered). Reports device
    microseconds per exchange and effective NVLink GB/s per rank."""
    from alpro_b200 import comm as acomm
    D = Dist()
    device, rank, world = D.device, D.rank, D.world
    comm = acomm.TorchDistComm() if world > 1 else None
    rows = []
    for G in (256, 512, 1024):
        if G % world:
            continue
        b = G // world
        x = torch.randn(b, 512, device=device)
        dg = torch.randn(G, 512, device=device)

        def exchange():
            if comm is None:
                return x, dg[:b]
            g_ = comm.all_gather(x)
            return g_, comm.reduce_scatter_sum(dg)
        for _ in range(10):
            exchange()
        n = 200
        ms = D.timed(exchange, n)
        us = ms * 1e3
        # bytes a rank moves over NVLink per exchange: receives (W-1)/W of the gathered buffer, and the same amount
        # flows the other way in the reduce-scatter
        nbytes = 2 * G * 512 * 4 * (world - 1) / max(world, 1)
        rows.append({"global_batch": G, "rows_per_rank": b, "us_per_exchange": round(us, 2),
                     "nvlink_bytes_per_rank": int(nbytes), "effective_gb_s": round(nbytes / (us * 1e-6) / 1e9, 2)})
    if rank == 0:
        res = {"metric": "VTC feature exchange latency (all-gather [B,512] fp32 + reduce-scatter of its gradient)",
               "value": rows[-1]["us_per_exchange"] if rows else None, "unit": "us", "n_gpus": world, "steps": 200,
               "warmup": 10, "ms_per_step": rows[-1]["us_per_exchange"] / 1e3 if rows else None,
               "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "impl": "alpro_b200", "config": {"workload": "vtc_allgather", "global_batches": [r["global_batch"] for r in rows],
                                              "parallelism": f"dp{world}"},
               "table": rows, "nvlink_peak_gb_s_per_dir": 900,
               "note": "latency-bound (<= 2 MB per exchange): the figure that matters is microseconds, not GB/s"}
        print(json.dumps(res), flush=True)
    D.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="pretrain",
                    choices=["pretrain", "retrieval", "timesformer_fwd", "vtc_allgather"])
    ap.add_argument("--batch", type=int, default=32, help="clips per GPU")
    ap.add_argument("--dtype", default=None, choices=["fp16", "bf16"],
                    help="GEMM operand format (default fp16; bf16 for --workload timesformer_fwd as BASELINE names it)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-torch-baseline", action="store_true")
    ap.add_argument("--no-dp-parity", action="store_true")
    ap.add_argument("--no-optimizer", action="store_true", help="skip the extra fwd+bwd+FusedAdamW timing")
    ap.add_argument("--ncu", action="store_true", help="one warm step, then one step between cudaProfilerStart/Stop")
    args = ap.parse_args()
    if args.dtype is None:
        args.dtype = "bf16" if args.workload == "timesformer_fwd" else "fp16"
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun when invoked directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", __file__] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.workload == "timesformer_fwd":
        run_timesformer_fwd(args)
    elif args.workload == "vtc_allgather":
        run_vtc_allgather(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
