#!/bin/bash
# 2-GPU call: N-rank == 1-rank bit-for-bit over NCCL, and the N = 2 bench line
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_multigpu.py -x -q -s > gpurun_out/r2g_multigpu.log 2>&1; echo "multigpu rc=$?"; tail -5 gpurun_out/r2g_multigpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2g_bench_N2.json 2> gpurun_out/r2g_bench_N2.err; echo "bench N2 rc=$?"
tail -c 400 gpurun_out/r2g_bench_N2.err
python -c "
import json; d=json.loads([l for l in open("gpurun_out/r2g_bench_N2.json") if l.startswith("{")][-1]); print(d['n_gpus'], d['value'], d['e2e'], d['ms_per_step'])"
