#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2e_bench.json')); print(d['value'], d['e2e'], d['unet_ms_per_ddim_step'], d['decode'], d['gpu_launches'])"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_ncu_forward_launches.csv python tests/gpu_ncu_forward.py > gpurun_out/r2e_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python tests/ncu_summarize.py gpurun_out/r2_ncu_forward_launches.csv gpurun_out/r2_ncu_forward_launches_summary "UNet forward N=8 64x128 (round 2 build)" | tail -25
for k in geglu lin320 conv attn; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_conv_kernel|attention_kernel" -s 2 -c 1 -o gpurun_out/r2_full_$k -f python tests/gpu_ncu_attn.py $k > gpurun_out/r2e_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
ls -la gpurun_out/*.ncu-rep
