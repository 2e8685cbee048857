#!/bin/bash
# vectorised im2col of the first conv: UNet parity (reference goldens) + forward timing + kernel time
mkdir -p gpurun_out
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum -k regex:im2col --clock-control none --csv --log-file gpurun_out/r2y_im2col.csv python tests/gpu_ncu_forward.py > /dev/null 2>&1; grep -o '"gpu__time_duration.sum","ns","[0-9]*"' gpurun_out/r2y_im2col.csv
echo "$(timeout 300 python tests/gpu_time_forward.py 40 2>&1 | tail -1)"
timeout 1500 python -m pytest tests/test_unet_gpu.py tests/test_vae_gpu.py -m gpu -x -q 2>&1 | tail -2
