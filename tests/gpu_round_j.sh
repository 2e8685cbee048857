#!/bin/bash
mkdir -p gpurun_out
LR_CASE_TIMEOUT=90 timeout 900 python tests/gpu_diag_ops.py --only linear,conv_mid,conv_2sm,conv_halo,conv_odd,gnf_linear,gnf_conv_1cta,upconv_small > gpurun_out/r2j_diag.log 2>&1; tail -1 gpurun_out/r2j_diag.log; grep "FAIL\|TIMEOUT" gpurun_out/r2j_diag.log | head
for i in 1 2; do
timeout 300 python tests/gpu_time_forward.py 40 2>&1 | tail -2
LR_GEMM_DEBUG=256 timeout 300 python tests/gpu_time_forward.py 40 2>&1 | tail -2
done
timeout 300 python tests/gpu_profile_steps.py gpurun_out/r2j_steps.txt > gpurun_out/r2j_steps.log 2>&1; head -3 gpurun_out/r2j_steps.txt; grep "+res bn\|c=320+0->320 bn\|c=320+0->960" gpurun_out/r2j_steps.txt | head -12
