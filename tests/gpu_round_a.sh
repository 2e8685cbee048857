#!/bin/bash
# round-2 GPU call A: full GPU test suite, op diagnostics (adversarial cases), tolerance table, bench (c2), step profile
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 400 python tests/gpu_tolerance_table.py gpurun_out/r2a_tolerance.md > gpurun_out/r2a_tolerance.log 2>&1; echo "tol rc=$?"
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2a_bench.json
timeout 300 python tests/gpu_profile_steps.py gpurun_out/r2a_steps.txt > gpurun_out/r2a_steps.log 2>&1; echo "steps rc=$?"
