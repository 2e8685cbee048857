#!/bin/bash
# 1-GPU call H: full GPU suite, c4 / c5 bench lines, compute-sanitizer (memcheck + racecheck) on small op cases
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2h_pytest.log
timeout 900 python bench.py --config c5 --steps 3 --warmup 3 --no-gpu-reference > gpurun_out/r2h_bench_c5.json 2> gpurun_out/r2h_bench_c5.err; echo "c5 rc=$?"
timeout 1200 python bench.py --config c4 --steps 2 --warmup 3 --no-gpu-reference > gpurun_out/r2h_bench_c4.json 2> gpurun_out/r2h_bench_c4.err; echo "c4 rc=$?"; tail -3 gpurun_out/r2h_bench_c4.err
python - <<'PY'
import json
for c in ("c5", "c4"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r2h_bench_{c}.json") if l.startswith("{")][-1])
        print(c, d["value"], d["e2e"]["value"], d["unet_ms_per_ddim_step"], d["decode"]["ms_per_batch"] if d["decode"] else None)
    except Exception as e:
        print(c, "parse failed", e)
PY
for tool in memcheck racecheck; do
  for c in linear_2sm_160 conv_halo_ragged conv_splitk attn_ragged gnf_conv gnf_linear upconv_small gn ln; do
    timeout 300 compute-sanitizer --tool $tool --error-exitcode 9 python tests/gpu_diag_ops.py --case $c > gpurun_out/r2h_san_${tool}_$c.log 2>&1
    echo "$tool $c rc=$? $(grep -c 'ERROR SUMMARY: 0 errors' gpurun_out/r2h_san_${tool}_$c.log) $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/r2h_san_${tool}_$c.log | tail -1)"
  done
done 2>&1 | tee gpurun_out/r2h_sanitizer_summary.txt
