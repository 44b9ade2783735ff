"""Developer probe: time both forward attention implementations at bench scale, then run a tiny model step with the
tcgen05 path under a watchdog that dumps the Python stack if a launch never returns."""
import faulthandler, os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from alpro_b200 import ops

faulthandler.enable()
dev = "cuda"


def bench(S, nseq, heads, vit, tc):
    os.environ["ALPRO_ATTN_TC"] = "1" if tc else "0"
    d = heads * 64
    g = torch.Generator(device=dev).manual_seed(1)
    if vit:
        T = 8; B = nseq // T; N = S - 1; Sc = 1 + N * T
        qkv = torch.randn(B * Sc, 3 * d, device=dev, generator=g).half()
        o = torch.zeros(B * Sc, d, device=dev, dtype=torch.float16)
        cls_o = torch.empty(nseq, d, device=dev, dtype=torch.float16)
        lse = torch.empty(nseq, heads, S, device=dev)
        f = lambda: ops.seq_attn_fwd(qkv, None, o, cls_o, lse, S, nseq, heads, T, T, Sc, 0.125)
    else:
        qkv = torch.randn(nseq * S, 3 * d, device=dev, generator=g).half()
        mask = torch.zeros(nseq, S, device=dev)
        o = torch.zeros(nseq * S, d, device=dev, dtype=torch.float16)
        lse = torch.empty(nseq, heads, S, device=dev)
        f = lambda: ops.seq_attn_fwd(qkv, mask, o, None, lse, S, nseq, heads, 1, 1, S, 0.125, 0.1, 1234)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    fl = 4.0 * nseq * heads * S * S * 64
    print(f"S={S} nseq={nseq} heads={heads} vit={vit} tc={tc}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s", flush=True)
    return o.float().clone()


for S, nseq, vit in ((197, 256, True), (40, 32, False), (237, 32, False), (237, 64, False)):
    a = bench(S, nseq, 12, vit, False)
    b = bench(S, nseq, 12, vit, True)
    print("   max |diff| =", float((a - b).abs().max()), flush=True)

print("model step with tcgen05 attention", flush=True)
os.environ["ALPRO_ATTN_TC"] = "1"
os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
faulthandler.dump_traceback_later(40, exit=True)
from oracle import configs
from tests import helpers
from tests.test_gpu_parity import build_cuda_model, to_cuda
cfg = configs.GOLDEN["tiny_retrieval"]
spec, sd, batch = helpers.make_inputs(cfg)
model = build_cuda_model(cfg, sd)
out = model(to_cuda(batch))
torch.cuda.synchronize()
print("forward ok", float(out["itc_loss"]), flush=True)
(out["itc_loss"] + out["itm_loss"]).backward()
torch.cuda.synchronize()
print("backward ok", flush=True)
