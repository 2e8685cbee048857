#!/bin/bash
# final round-2 validation: full GPU suite, bench lines (c2 default / c5 / c4), ncu launch list + full captures of the
# reworked kernels, per-step profile, compute-sanitizer on the new code paths, smoke
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2s_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2s_bench_N1.json 2> gpurun_out/r2s_bench_N1.err; echo "bench rc=$?"; tail -2 gpurun_out/r2s_bench_N1.err
timeout 900 python bench.py --config c5 --steps 3 --warmup 3 --no-gpu-reference > gpurun_out/r2s_bench_c5.json 2> gpurun_out/r2s_bench_c5.err; echo "c5 rc=$?"
timeout 1200 python bench.py --config c4 --steps 2 --warmup 3 --no-gpu-reference > gpurun_out/r2s_bench_c4.json 2> gpurun_out/r2s_bench_c4.err; echo "c4 rc=$?"
python - <<'PY'
import json
for c in ("N1", "c5", "c4"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r2s_bench_{c}.json") if l.startswith("{")][-1])
        print(c, d["value"], d["e2e"]["value"], d["unet_ms_per_ddim_step"], d["decode"]["ms_per_batch"] if d.get("decode") else None,
              d["roofline"]["frac"], d.get("gpu_reference"), d.get("cpu_baseline", {}) and d["cpu_baseline"].get("value"))
    except Exception as e:
        print(c, "parse failed", e)
PY
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_ncu_forward_launches.csv python tests/gpu_ncu_forward.py > gpurun_out/r2s_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python tests/ncu_summarize.py gpurun_out/r2_ncu_forward_launches.csv gpurun_out/r2_ncu_forward_launches_summary "UNet forward N=8 64x128 (round 2 final build)" | tail -22
for k in geglu lin320; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_conv_kernel" -s 2 -c 1 -o gpurun_out/r2s_full_$k -f python tests/gpu_ncu_attn.py $k > gpurun_out/r2s_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gn_persistent_kernel" -s 2 -c 1 -o gpurun_out/r2s_full_gn -f python tests/gpu_ncu_attn.py gn > gpurun_out/r2s_ncu_gn.log 2>&1; echo "ncu gn rc=$?"
timeout 300 python tests/gpu_profile_steps.py gpurun_out/r2s_steps.txt > gpurun_out/r2s_steps.log 2>&1; head -3 gpurun_out/r2s_steps.txt
for c in gn_bigmean_twopass linear_2sm_160 linear_2sm_geglu linear_ragged attn_cross; do
  LR_ATTN_PERSIST=1 timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/gpu_diag_ops.py --case $c > gpurun_out/r2s_san_memcheck_$c.log 2>&1
  echo "memcheck $c rc=$? $(grep 'ERROR SUMMARY' gpurun_out/r2s_san_memcheck_$c.log | tail -1) $(grep -c PASS gpurun_out/r2s_san_memcheck_$c.log)"
done 2>&1 | tee gpurun_out/r2s_sanitizer_summary.txt
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
