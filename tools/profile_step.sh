#!/bin/bash
# ncu evidence for one training step (run under gpurun, 1 GPU). Outputs land in gpurun_out/.
#   1) launch list: every kernel of one step with its device time (cold-cache, serialised: compare SHARES)
#   2) full capture of the dominant kernel (gemm16) for three launches
set -x
B=${1:-32}
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python bench.py --ncu --batch $B > gpurun_out/ncu_launch.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm16 -s 60 -c 6 \
    -o gpurun_out/prof_gemm_step python bench.py --ncu --batch $B > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
