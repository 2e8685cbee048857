#!/bin/bash
# attention: persistent kernel for short key sequences only; parity + same-call A/B against the committed build
mkdir -p gpurun_out
LR_CASE_TIMEOUT=90 timeout 600 python tests/gpu_diag_ops.py --only attn > gpurun_out/r2r_diag.log 2>&1; tail -1 gpurun_out/r2r_diag.log
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -x -q 2>&1 | tail -2
for rep in 1 2 3; do
for v in head cur; do
  lib=$PWD/leftrefill_b200/ab/liblr_$v.so; [ $v = cur ] && lib=$PWD/leftrefill_b200/liblr_b200.so
  r=$(LR_B200_LIB=$lib timeout 300 python tests/gpu_time_forward.py 40 2>&1 | tail -1)
  echo "$v: $r"
done
done
