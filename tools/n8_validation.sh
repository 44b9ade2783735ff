#!/bin/bash
# gpurun --gpus N -- bash tools/n8_validation.sh N [tag]    scaling bench line with dp_parity (default copy-engine
# reducer) and, for comparison, the NCCL reducer on the same box. A healthy run takes ~90 s per line; the timeouts are
# short on purpose: a hung collective at N GPUs burns N x the box time (a rank-asymmetric loop in bench.py once cost a
# whole round's remaining GPU budget here).
N=${1:-8}
TAG=${2:-r02}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29633"
timeout 300 $TR bench.py --gpus $N --steps 8 --warmup 3 --no-optimizer > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
ALPRO_GRAD_REDUCER=nccl timeout 300 $TR bench.py --gpus $N --steps 8 --warmup 3 --no-optimizer --no-dp-parity > gpurun_out/${TAG}_bench_n${N}_nccl.json 2> gpurun_out/${TAG}_bench_n${N}_nccl.err
for f in gpurun_out/${TAG}_bench_n$N*.json; do echo "== $f"; grep '^{' $f | cut -c1-200; done
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench_n$N.json", "gpurun_out/${TAG}_bench_n${N}_nccl.json"):
    for l in open(f):
        if l.startswith("{"):
            r = json.loads(l)
            print(f, r["value"], r["ms_per_step"], r["e2e"]["value"], r.get("dp_parity"), r["clocks"])
PY
tail -3 gpurun_out/${TAG}_bench_n$N.err
