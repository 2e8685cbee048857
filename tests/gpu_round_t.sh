#!/bin/bash
# GEGLU: (v, v', g, g') row order + packed fp32 GELU in the lean epilogue: parity, same-call A/B against the previous build
mkdir -p gpurun_out
LR_CASE_TIMEOUT=90 timeout 600 python tests/gpu_diag_ops.py --only geglu,linear > gpurun_out/r2t_diag.log 2>&1; tail -1 gpurun_out/r2t_diag.log; grep "FAIL\|TIMEOUT" gpurun_out/r2t_diag.log | head
for rep in 1 2 3; do
for v in prev cur; do
  lib=$PWD/leftrefill_b200/ab/liblr_$v.so; [ $v = cur ] && lib=$PWD/leftrefill_b200/liblr_b200.so
  r=$(LR_B200_LIB=$lib timeout 300 python tests/gpu_time_forward.py 40 2>&1 | tail -1)
  echo "$v: $r"
done
done
timeout 300 python tests/gpu_profile_steps.py gpurun_out/r2t_steps.txt > /dev/null 2>&1; grep geglu gpurun_out/r2t_steps.txt
timeout 1800 python -m pytest tests/test_unet_gpu.py tests/test_ops_gpu.py -m gpu -x -q 2>&1 | tail -3
