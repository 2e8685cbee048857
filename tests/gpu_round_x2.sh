#!/bin/bash
# 2-GPU call: NCCL 2 ranks == 1 rank bit for bit, the single-GPU sharding test, lean == general epilogue
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_ops_gpu.py -m gpu -x -q -k "sharding or lean_and_general or geglu" 2>&1 | tail -3
