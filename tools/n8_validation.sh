#!/bin/bash
# gpurun --gpus 8 -- bash tools/n8_validation.sh     one box acquisition: scaling bench line with dp_parity (default
# copy-engine reducer), the NCCL reducer for comparison, retrieval at 8 GPUs, VTC exchange microbench.
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29633"
timeout 600 $TR bench.py --gpus $N --steps 8 --warmup 3 --no-optimizer > gpurun_out/r02f_bench_n$N.json 2> gpurun_out/r02f_bench_n$N.err
ALPRO_GRAD_REDUCER=nccl timeout 600 $TR bench.py --gpus $N --steps 8 --warmup 3 --no-optimizer --no-dp-parity > gpurun_out/r02f_bench_n${N}_nccl.json 2> gpurun_out/r02f_bench_n${N}_nccl.err
timeout 600 $TR bench.py --gpus $N --workload retrieval --steps 8 --warmup 3 --no-optimizer --no-dp-parity > gpurun_out/r02f_retrieval_n$N.json 2> gpurun_out/r02f_retrieval_n$N.err
timeout 300 $TR bench.py --gpus $N --workload vtc_allgather > gpurun_out/r02f_vtc_n$N.json 2> gpurun_out/r02f_vtc_n$N.err
for f in gpurun_out/r02f_*_n$N*.json; do echo "== $f"; cut -c1-1200 $f; done
tail -3 gpurun_out/r02f_bench_n$N.err
