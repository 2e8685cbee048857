#!/bin/bash
# last check of the shipped build: full GPU suite, tolerance table, smoke, a short bench run
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2z_pytest.log
timeout 900 python tests/gpu_tolerance_table.py > gpurun_out/r2z_tolerance_table.md 2> gpurun_out/r2z_tolerance.err; echo "tolerance rc=$?"; tail -12 gpurun_out/r2z_tolerance_table.md
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 2 --warmup 3 --no-gpu-reference > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2z_bench.json') if l.startswith('{')][-1]); print(d['value'], d['e2e']['value'], d['unet_ms_per_ddim_step'], d['roofline']['frac'], d['roofline']['traffic'], d['gpu_launches'], d['clocks'])"
