#!/bin/bash
# lean epilogue + distributed store issue + producer-side residual fetch: parity, traces, forward timing A/B
mkdir -p gpurun_out
LR_CASE_TIMEOUT=90 timeout 900 python tests/gpu_diag_ops.py > gpurun_out/r2l_diag.log 2>&1; tail -1 gpurun_out/r2l_diag.log; grep "FAIL\|TIMEOUT" gpurun_out/r2l_diag.log | head
timeout 250 python tests/gpu_trace_lin.py > gpurun_out/r2l_trace_lean3.txt 2>&1
for i in 1 2; do
timeout 300 python tests/gpu_time_forward.py 40 2>&1 | tail -1
LR_NO_LEAN_EPI=1 timeout 300 python tests/gpu_time_forward.py 40 2>&1 | tail -1
done
timeout 300 python tests/gpu_profile_steps.py gpurun_out/r2l_steps.txt > gpurun_out/r2l_steps.log 2>&1; head -30 gpurun_out/r2l_steps.txt
timeout 1200 python -m pytest tests/test_ops_gpu.py -m gpu -x -q 2>&1 | tail -5
