"""One UNet forward (N = 8, 64x128: the bench workload) bracketed by cudaProfilerStart/Stop for an ncu launch list:
    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/launches.csv python tests/gpu_ncu_forward.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from helpers import O, synthetic_inputs  # noqa: E402

import bench  # noqa: E402
import leftrefill_b200 as lr  # noqa: E402

dev = torch.device("cuda")
m = bench.device_unet(lr.UNetModel, O.DEFAULT_CFG, dev, seed=0)
xT, c_cat, ctx, uc = synthetic_inputs(4, device=dev)
xc = torch.cat([torch.cat([xT, xT]), torch.cat([c_cat, c_cat])], dim=1).contiguous()
tt = torch.full((8,), 981, dtype=torch.long, device=dev)
m.sync_weights()
m.set_context(torch.cat([uc, ctx]).contiguous())
for _ in range(2):
    m.forward_native(xc, tt, None)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
m.forward_native(xc, tt, None)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
