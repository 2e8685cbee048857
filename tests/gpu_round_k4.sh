#!/bin/bash
# 4-GPU call: BASELINE.json configs[4] (NVS, 16 canvases over 4 GPUs) and configs[2]-style c2 on 4 GPUs
mkdir -p gpurun_out
nvidia-smi -L | head -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --config c5 --steps 3 --warmup 3 > gpurun_out/r2k_bench_c5_N4.json 2> gpurun_out/r2k_bench_c5_N4.err; echo "c5 N4 rc=$?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2k_bench_c5_N4.json') if l.startswith('{')][-1]); print(d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['name'])"
