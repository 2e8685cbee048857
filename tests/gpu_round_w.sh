#!/bin/bash
# final check of the shipped build: full GPU suite, default bench line, smoke
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2w_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2w_bench_N1.json 2> gpurun_out/r2w_bench_N1.err; echo "bench rc=$?"; tail -2 gpurun_out/r2w_bench_N1.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2w_bench_ref.json 2> gpurun_out/r2w_bench_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/r2w_bench_ref.json
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2w_bench_N1.json") if l.startswith("{")][-1])
print("N1", d["value"], d["e2e"]["value"], d["unet_ms_per_ddim_step"], d["decode"]["ms_per_batch"], d["roofline"]["frac"], d["roofline"]["traffic"], d["clocks"])
PY
timeout 300 python tests/gpu_profile_steps.py gpurun_out/r2w_steps.txt > gpurun_out/r2w_steps.log 2>&1; head -3 gpurun_out/r2w_steps.txt
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
