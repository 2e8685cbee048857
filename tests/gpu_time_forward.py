"""Whole-forward timing (no per-step events) at the bench workload: N = 8 forward and the CFG-pair forward.
    python tests/gpu_time_forward.py [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import O, synthetic_inputs  # noqa: E402

import leftrefill_b200 as lr  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 30
cfg = O.DEFAULT_CFG
m = lr.UNetModel(**cfg)
m.load_state_dict(O.make_state_dict(cfg, seed=0), strict=True)
m = m.cuda().eval()
xT, c_cat, ctx, uc = synthetic_inputs(4, device="cuda")
xc = torch.cat([torch.cat([xT, xT]), torch.cat([c_cat, c_cat])], dim=1).contiguous()
tt = torch.full((8,), 981, dtype=torch.long, device="cuda")
m.sync_weights()
m.set_context(torch.cat([uc, ctx]).contiguous())


def timed(fn):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


print(f"LR_NO_PDL={os.environ.get('LR_NO_PDL')}  forward N=8: {timed(lambda: m.forward_native(xc, tt, None)):.3f} ms")
x4 = torch.cat([xT, c_cat], dim=1).contiguous()
t4 = tt[:4].contiguous()
if hasattr(m, "forward_native_cfg_pair"):
    try:
        print(f"cfg-pair forward (4 canvases): {timed(lambda: m.forward_native_cfg_pair(x4, t4)):.3f} ms")
    except Exception as e:  # signature differences are not this script's concern
        print("cfg-pair timing skipped:", e)
