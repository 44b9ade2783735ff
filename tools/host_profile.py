"""Where does the HOST time of one training step go? cProfile over a few un-synchronised steps of the bench workload
(run on the GPU box): python tools/host_profile.py [batch]"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
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
t0 = time.perf_counter()
for _ in range(3):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0) / 3:.1f} ms/step, wall {1e3 * (t2 - t0) / 3:.1f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(32)
st.sort_stats("cumtime").print_stats(28)
