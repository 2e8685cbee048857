#!/bin/bash
# LayerNorm row statistics from the producing Linear's (lean) epilogue: A/B + parity of the engine variants
mkdir -p gpurun_out
for i in 1 2; do
timeout 300 python tests/gpu_time_forward.py 40 2>&1 | tail -1
LR_LN_ROWSTATS=1 timeout 300 python tests/gpu_time_forward.py 40 2>&1 | tail -1
done
LR_LN_ROWSTATS=1 timeout 300 python tests/gpu_profile_steps.py gpurun_out/r2p_steps_rowstats.txt > gpurun_out/r2p_steps.log 2>&1; head -3 gpurun_out/r2p_steps_rowstats.txt; grep "linear" gpurun_out/r2p_steps_rowstats.txt | head -16
timeout 1200 python -m pytest tests/test_unet_gpu.py -m gpu -x -q -k "variants" 2>&1 | tail -3
