#!/bin/bash
# small_linear_kernel (timestep path): vectorised staging, one output per warp for the narrow layers
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum -k regex:small_linear --clock-control none --csv --log-file gpurun_out/r2y_small_linear.csv python tests/gpu_ncu_forward.py > /dev/null 2>&1; grep -o '"gpu__time_duration.sum","ns","[0-9]*"' gpurun_out/r2y_small_linear.csv
for rep in 1 2; do echo "$(timeout 300 python tests/gpu_time_forward.py 40 2>&1 | tail -1)"; done
timeout 1800 python -m pytest tests/test_unet_gpu.py -m gpu -x -q 2>&1 | tail -2
