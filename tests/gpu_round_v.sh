#!/bin/bash
# ln_stats_kernel at 3 CTAs per SM (80 registers): same-call A/B
mkdir -p gpurun_out
for v in prev cur; do
  lib=$PWD/leftrefill_b200/ab/liblr_$v.so; [ $v = cur ] && lib=$PWD/leftrefill_b200/liblr_b200.so
  LR_B200_LIB=$lib timeout 300 python tests/gpu_profile_steps.py gpurun_out/r2v_steps_$v.txt > /dev/null 2>&1; echo "== $v"; grep "layernorm" gpurun_out/r2v_steps_$v.txt
done
for rep in 1 2; do
for v in prev cur; do
  lib=$PWD/leftrefill_b200/ab/liblr_$v.so; [ $v = cur ] && lib=$PWD/leftrefill_b200/liblr_b200.so
  r=$(LR_B200_LIB=$lib timeout 300 python tests/gpu_time_forward.py 40 2>&1 | tail -1)
  echo "$v: $r"
done
done
