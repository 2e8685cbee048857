#!/bin/bash
# persistent attention kernel: parity (op-level diag + pytest), A/B forward timing, per-step profile
mkdir -p gpurun_out
LR_CASE_TIMEOUT=90 timeout 600 python tests/gpu_diag_ops.py --only attn > gpurun_out/r2q_diag.log 2>&1; tail -1 gpurun_out/r2q_diag.log; grep "FAIL\|TIMEOUT" gpurun_out/r2q_diag.log | head
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "attn or attention" 2>&1 | tail -2
for i in 1 2; do
timeout 300 python tests/gpu_time_forward.py 40 2>&1 | tail -1
LR_ATTN_NO_PERSIST=1 timeout 300 python tests/gpu_time_forward.py 40 2>&1 | tail -1
done
timeout 300 python tests/gpu_profile_steps.py gpurun_out/r2q_steps.txt > gpurun_out/r2q_steps.log 2>&1; head -3 gpurun_out/r2q_steps.txt; grep attention gpurun_out/r2q_steps.txt
