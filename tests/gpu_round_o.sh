#!/bin/bash
# persistent GroupNorm (37 chunks / image, packed fp32 apply): norm timings, phase trace, full GPU suite, forward + steps
mkdir -p gpurun_out
timeout 600 python tests/gpu_time_norm.py > gpurun_out/r2o_norm.txt 2>&1; grep "default" gpurun_out/r2o_norm.txt
timeout 300 python tests/gpu_gn_trace.py > gpurun_out/r2o_gn_trace.txt 2>&1
for i in 1 2; do timeout 300 python tests/gpu_time_forward.py 40 2>&1 | tail -1; done
timeout 300 python tests/gpu_profile_steps.py gpurun_out/r2o_steps.txt > gpurun_out/r2o_steps.log 2>&1; head -3 gpurun_out/r2o_steps.txt; grep groupnorm gpurun_out/r2o_steps.txt | head -12
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
