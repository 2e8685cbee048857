#!/bin/bash
# lean epilogue incl. folded LayerNorm: full GPU suite, forward timing A/B, per-step profile
mkdir -p gpurun_out
for i in 1 2; do
timeout 300 python tests/gpu_time_forward.py 40 2>&1 | tail -1
LR_NO_LEAN_EPI=1 timeout 300 python tests/gpu_time_forward.py 40 2>&1 | tail -1
done
timeout 300 python tests/gpu_profile_steps.py gpurun_out/r2m_steps.txt > gpurun_out/r2m_steps.log 2>&1; head -24 gpurun_out/r2m_steps.txt
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
