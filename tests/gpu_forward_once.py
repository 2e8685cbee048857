"""One profiled UNet forward at the bench workload (N = 8, 64x128) for ncu:
    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file launches.csv python tests/gpu_forward_once.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import O, synthetic_inputs  # noqa: E402

import leftrefill_b200 as lr  # noqa: E402
from leftrefill_b200 import _native as N  # noqa: E402

cfg = O.DEFAULT_CFG
m = lr.UNetModel(**cfg)
m.load_state_dict(O.make_state_dict(cfg, seed=0), strict=True)
m = m.cuda().eval()
xT, c_cat, ctx, uc = synthetic_inputs(4, device="cuda")
xc = torch.cat([torch.cat([xT, xT]), torch.cat([c_cat, c_cat])], dim=1).contiguous()
tt = torch.full((8,), 981, dtype=torch.long, device="cuda")
m.sync_weights()
m.set_context(torch.cat([uc, ctx]).contiguous())
for _ in range(2):
    m.forward_native(xc, tt, None)
torch.cuda.synchronize()
N.lib().lr_launch_count_reset()
torch.cuda.profiler.start()
m.forward_native(xc, tt, None)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches in the profiled forward:", N.lib().lr_launch_count())
