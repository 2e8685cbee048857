#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_unet_gpu.py -x -q -k "variants or full_size_properties or cfg_pair" > gpurun_out/r2i_pytest_variants.log 2>&1; echo "variants rc=$?"; tail -4 gpurun_out/r2i_pytest_variants.log
for i in 1 2; do
timeout 300 python tests/gpu_time_forward.py 40 2>&1 | tail -2
LR_NO_LN_ROWSTATS=1 timeout 300 python tests/gpu_time_forward.py 40 2>&1 | tail -2
done
timeout 300 python tests/gpu_profile_steps.py gpurun_out/r2i_steps.txt > gpurun_out/r2i_steps.log 2>&1; head -5 gpurun_out/r2i_steps.txt; grep "layernorm\|+res bn" gpurun_out/r2i_steps.txt | head -12
for c in conv_halo_1cta linear_bn160 gnf_conv_1cta gnf_linear_1cta; do
  timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python tests/gpu_diag_ops.py --case $c > gpurun_out/r2i_san_racecheck_$c.log 2>&1
  echo "racecheck(1-CTA build of the same kernel) $c rc=$? $(grep 'RACECHECK SUMMARY' gpurun_out/r2i_san_racecheck_$c.log | tail -1)"
done 2>&1 | tee gpurun_out/r2i_sanitizer_1cta.txt
