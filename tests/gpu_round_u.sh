#!/bin/bash
# GroupNorm SiLU: one reciprocal per pair on the FMA pipe (Newton) vs both on the MUFU: parity, timings, same-call A/B
mkdir -p gpurun_out
LR_CASE_TIMEOUT=90 timeout 600 python tests/gpu_diag_ops.py --only gn,groupnorm > gpurun_out/r2u_diag.log 2>&1; tail -1 gpurun_out/r2u_diag.log; grep "FAIL\|TIMEOUT" gpurun_out/r2u_diag.log | head
for v in prev cur; do
  lib=$PWD/leftrefill_b200/ab/liblr_$v.so; [ $v = cur ] && lib=$PWD/leftrefill_b200/liblr_b200.so
  echo "== $v"; LR_B200_LIB=$lib LR_GN_SHAPES=3 LR_GN_PASSES="0" timeout 300 python tests/gpu_time_gn_passes.py
  LR_B200_LIB=$lib timeout 300 python tests/gpu_gn_trace.py | grep "cta 0" | sed -n '2p;5p;8p'
done
for rep in 1 2 3; do
for v in prev cur; do
  lib=$PWD/leftrefill_b200/ab/liblr_$v.so; [ $v = cur ] && lib=$PWD/leftrefill_b200/liblr_b200.so
  r=$(LR_B200_LIB=$lib timeout 300 python tests/gpu_time_forward.py 40 2>&1 | tail -1)
  echo "$v: $r"
done
done
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -x -q 2>&1 | tail -2
