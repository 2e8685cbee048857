#!/bin/bash
# persistent GroupNorm with two CTAs per SM (64 registers, batches of 4) vs one CTA per SM (128 registers, batches of 8)
mkdir -p gpurun_out
LR_CASE_TIMEOUT=90 timeout 600 python tests/gpu_diag_ops.py --only gn,groupnorm > gpurun_out/r2u_diag.log 2>&1; tail -1 gpurun_out/r2u_diag.log; grep "FAIL\|TIMEOUT" gpurun_out/r2u_diag.log | head
for v in prev cur; do
  lib=$PWD/leftrefill_b200/ab/liblr_$v.so; [ $v = cur ] && lib=$PWD/leftrefill_b200/liblr_b200.so
  echo "== $v"; LR_B200_LIB=$lib timeout 300 python tests/gpu_time_norm.py run 2>&1 | grep GN
done
for rep in 1 2 3; do
for v in prev cur; do
  lib=$PWD/leftrefill_b200/ab/liblr_$v.so; [ $v = cur ] && lib=$PWD/leftrefill_b200/liblr_b200.so
  r=$(LR_B200_LIB=$lib timeout 300 python tests/gpu_time_forward.py 40 2>&1 | tail -1)
  echo "$v: $r"
done
done
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_vae_gpu.py -m gpu -x -q 2>&1 | tail -2
