#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tests/gpu_time_gnfuse.py > gpurun_out/r2c_gnfuse.log 2>&1; echo "gnfuse rc=$?"
cat gpurun_out/r2c_gnfuse.log
LR_CASE_TIMEOUT=90 timeout 600 python tests/gpu_diag_ops.py --only gnf > gpurun_out/r2c_diag.log 2>&1; tail -2 gpurun_out/r2c_diag.log
timeout 300 python tests/gpu_profile_steps.py gpurun_out/r2c_steps.txt > gpurun_out/r2c_steps.log 2>&1; echo "steps rc=$?"
head -24 gpurun_out/r2c_steps.txt
LR_GN_FUSE_CONV=1 timeout 300 python tests/gpu_profile_steps.py gpurun_out/r2c_steps_fuseconv.txt > gpurun_out/r2c_steps2.log 2>&1; echo "steps2 rc=$?"
head -24 gpurun_out/r2c_steps_fuseconv.txt
timeout 300 python tests/gpu_time_forward.py 30 2>&1 | tail -3
LR_GN_FUSE_CONV=1 timeout 300 python tests/gpu_time_forward.py 30 2>&1 | tail -3
LR_NO_GN_FUSE=1 timeout 300 python tests/gpu_time_forward.py 30 2>&1 | tail -3
