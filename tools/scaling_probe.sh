#!/bin/bash
# gpurun --gpus N -- bash tools/scaling_probe.sh N [set]     (results: gpurun_out/scaling_probe_nN.jsonl)
N=${1:-2}
SET=${2:-nvl}
OUT=gpurun_out/scaling_probe_n$N.jsonl
mkdir -p gpurun_out; : > $OUT
run() {  # tag, env...
  tag=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
      tools/scaling_probe.py $tag 2>gpurun_out/probe_$tag.err | grep '^{' >> $OUT
}
if [ "$SET" = "nccl" ]; then
  python tools/scaling_probe.py single_gpu 2>/dev/null | grep '^{' >> $OUT
  run procs_only_no_comm PROBE_COMM=0
  run vtc_only_no_grad_reduce PROBE_NO_GRAD_REDUCE=1
  run overlap_fp32 ALPRO_GRAD_REDUCER=nccl ALPRO_DP_OVERLAP=1
  run end_of_step_fp32 ALPRO_DP_OVERLAP=0
  run overlap_fp32_maxctas8 ALPRO_GRAD_REDUCER=nccl ALPRO_DP_OVERLAP=1 NCCL_MAX_CTAS=8
  run overlap_bf16 ALPRO_DP_OVERLAP=1 ALPRO_GRAD_COMPRESS=bf16
  run overlap_bf16_maxctas8 ALPRO_DP_OVERLAP=1 ALPRO_GRAD_COMPRESS=bf16 NCCL_MAX_CTAS=8
  run end_of_step_bf16 ALPRO_DP_OVERLAP=stream ALPRO_GRAD_COMPRESS=bf16
else
  run vtc_only_no_grad_reduce PROBE_NO_GRAD_REDUCE=1
  run nccl_overlap ALPRO_GRAD_REDUCER=nccl
  run nccl_end_of_step ALPRO_GRAD_REDUCER=nccl ALPRO_DP_OVERLAP=0
  run ce ALPRO_GRAD_REDUCER=ce
  run nvl_ctas64 ALPRO_GRAD_REDUCER=nvl ALPRO_NVL_CTAS=64
  run nvl_ctas148 ALPRO_GRAD_REDUCER=nvl ALPRO_NVL_CTAS=148
fi
cat $OUT
