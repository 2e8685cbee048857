#!/bin/bash
# round-2 GPU call B: fused GroupNorm bring-up (op cases first, each in its own subprocess with a timeout), then the suite
mkdir -p gpurun_out
LR_CASE_TIMEOUT=90 timeout 900 python tests/gpu_diag_ops.py --only gnf,conv_halo,conv_mid,conv_2sm,linear_2sm,linear_ragged,linear_big,gn,attn_bigrange,geglu_big > gpurun_out/r2b_diag.log 2>&1
tail -3 gpurun_out/r2b_diag.log
grep -c "^   PASS" gpurun_out/r2b_diag.log; grep "FAIL\|TIMEOUT\|rc=[^0]" gpurun_out/r2b_diag.log | head -20
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_pytest.log
timeout 300 python tests/gpu_profile_steps.py gpurun_out/r2b_steps.txt > gpurun_out/r2b_steps.log 2>&1; echo "steps rc=$?"
head -12 gpurun_out/r2b_steps.txt
timeout 300 python tests/gpu_diag_determinism.py > gpurun_out/r2b_det.log 2>&1; echo "det rc=$?"; tail -5 gpurun_out/r2b_det.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2b_bench.json')); print(d['value'], d['e2e']['value'], d['unet_ms_per_ddim_step'], d['roofline']['by_class'])"
