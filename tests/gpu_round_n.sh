#!/bin/bash
# persistent GroupNorm kernel: parity, timings per scheme, forward timing, VAE + UNet suites
mkdir -p gpurun_out
LR_CASE_TIMEOUT=90 timeout 600 python tests/gpu_diag_ops.py --only groupnorm,gn > gpurun_out/r2n_diag.log 2>&1; tail -1 gpurun_out/r2n_diag.log; grep "FAIL\|TIMEOUT" gpurun_out/r2n_diag.log | head
timeout 600 python tests/gpu_time_norm.py > gpurun_out/r2n_norm.txt 2>&1; grep GN gpurun_out/r2n_norm.txt
for i in 1 2; do timeout 300 python tests/gpu_time_forward.py 40 2>&1 | tail -1; done
timeout 300 python tests/gpu_time_vae.py 2>&1 | tail -3
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
