"""Per-plan-step CUDA-event profile of one UNet forward at the bench workload (N = 8, 64x128). Writes a table grouped by
op shape: launches, total ms, TFLOP/s.   python tests/gpu_profile_steps.py [out.txt]"""
import ctypes
import os
import sys
from collections import OrderedDict

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import O, synthetic_inputs  # noqa: E402

import leftrefill_b200 as lr  # noqa: E402
from leftrefill_b200 import _native as N  # noqa: E402


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else None
    dev = "cuda"
    cfg = O.DEFAULT_CFG
    m = lr.UNetModel(**cfg)
    m.load_state_dict(O.make_state_dict(cfg, seed=0), strict=True)
    m = m.cuda().eval()
    xT, c_cat, ctx, uc = synthetic_inputs(4, device=dev)
    xc = torch.cat([torch.cat([xT, xT]), torch.cat([c_cat, c_cat])], dim=1).contiguous()
    tt = torch.full((8,), 981, dtype=torch.long, device=dev)
    m.sync_weights()
    m.set_context(torch.cat([uc, ctx]).contiguous())
    L, h = N.lib(), m.engine()
    for _ in range(3):
        m.forward_native(xc, tt, None)
    L.lr_unet_set_profiling(h, 1)
    iters = 5
    agg = OrderedDict()
    for it in range(iters + 1):
        m.forward_native(xc, tt, None)
        torch.cuda.synchronize()
        if it == 0:
            continue
        for i in range(L.lr_unet_num_steps(h)):
            ms, fl, cls = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
            buf = ctypes.create_string_buffer(256)
            N.check(L.lr_unet_step_info(h, i, ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(cls), buf, 256), "info")
            key = (cls.value, buf.value.decode() or "other")
            a = agg.setdefault(key, [0, 0.0, 0.0])
            a[0] += 1
            a[1] += ms.value
            a[2] += fl.value
    rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
    total = sum(v[1] for v in agg.values()) / iters
    lines = [f"UNet forward N=8 64x128, {L.lr_unet_num_steps(h)} plan steps, {total:.2f} ms/forward (event-timed, serialised)",
             f"{'ms/fwd':>8} {'share':>6} {'calls':>5} {'us/call':>8} {'TFLOP/s':>8}  op"]
    for (cls, desc), (cnt, ms, fl) in rows:
        per = ms / iters
        lines.append(f"{per:8.3f} {100 * per / total:5.1f}% {cnt // iters:5d} {1e3 * ms / cnt:8.1f} "
                     f"{(fl / ms / 1e9) if fl > 0 else 0:8.1f}  {desc}")
    txt = "\n".join(lines)
    print(txt)
    if out_path:
        open(out_path, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
