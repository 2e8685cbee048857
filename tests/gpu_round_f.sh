#!/bin/bash
mkdir -p gpurun_out
LR_CASE_TIMEOUT=90 timeout 600 python tests/gpu_diag_ops.py --only upconv,conv_small,conv_mid,linear_geglu,linear_2sm_geglu,conv_s2 > gpurun_out/r2f_diag.log 2>&1; grep "PASS\|FAIL\|rc=\|SUMMARY" gpurun_out/r2f_diag.log | tail -14
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2f_pytest.log
timeout 300 python tests/gpu_time_forward.py 30 2>&1 | tail -2
LR_NO_UPFOLD=1 timeout 300 python tests/gpu_time_forward.py 30 2>&1 | tail -2
timeout 300 python tests/gpu_time_vae.py 4 2>&1 | tail -2
LR_NO_UPFOLD=1 timeout 300 python tests/gpu_time_vae.py 4 2>&1 | tail -2
timeout 300 python tests/gpu_profile_steps.py gpurun_out/r2f_steps.txt > gpurun_out/r2f_steps.log 2>&1; grep "taps\|c=640+0->640 bn\|c=1280+0->1280 bn=.*tiles" gpurun_out/r2f_steps.txt | head
