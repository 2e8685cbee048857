#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vae_gpu.py -x -q > gpurun_out/r2d_vae_pytest.log 2>&1; echo "vae pytest rc=$?"; tail -15 gpurun_out/r2d_vae_pytest.log
timeout 300 python tests/gpu_time_vae.py 4 > gpurun_out/r2d_vae_time.log 2>&1; cat gpurun_out/r2d_vae_time.log | tail -4
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2d_pytest.log
timeout 300 python tests/gpu_time_forward.py 30 2>&1 | tail -2
