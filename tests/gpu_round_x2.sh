#!/bin/bash
# 2-GPU check of the last build: NCCL 2 ranks == 1 rank bit for bit, a short 2-rank bench run
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2x_bench_N2_last.json 2> gpurun_out/r2x_bench_N2_last.err; echo "N2 rc=$?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2x_bench_N2_last.json') if l.startswith('{')][-1]); print(d['n_gpus'], d['value'], d['e2e']['value'], d['unet_ms_per_ddim_step'], d['clocks'])"
