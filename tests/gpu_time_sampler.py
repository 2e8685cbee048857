"""Where does a DDIMSampler.sample call spend its time (bench workload: 4 canvases, 50 steps, cfg 2.5)?
Prints wall / GPU time per call with and without a device sync between calls, plus the GPU timeline of one call
(entry -> first step graph replay -> last replay -> return) taken with CUDA events inserted by monkey-patching."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import FakeLDM, O, synthetic_inputs  # noqa: E402

import leftrefill_b200 as lr  # noqa: E402
from leftrefill_b200 import ddim as D  # noqa: E402

dev = torch.device("cuda")
cfg = O.DEFAULT_CFG
unet = lr.UNetModel(**cfg)
unet.load_state_dict(O.make_state_dict(cfg, seed=0), strict=True)
unet = unet.to(dev).eval()
ldm = FakeLDM(unet, dev)
B, S, H, W = 4, 50, 64, 128
xT_h, ccat_h, ctx_h, uc_h = [t.pin_memory() for t in synthetic_inputs(B, h=H, w=W)]
xT, ccat, ctx, uc = [t.to(dev) for t in (xT_h, ccat_h, ctx_h, uc_h)]
marks = []
orig_replay = D._StepGraph.replay


def replay(self):
    if len(marks) < 2 * S + 4:
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        marks.append((time.perf_counter(), e))
    orig_replay(self)


D._StepGraph.replay = replay


def one(x, c, cx, u):
    s = lr.DDIMSampler(ldm)
    cond = {"c_concat": [c], "c_crossattn": [cx]}
    ucond = {"c_concat": [c], "c_crossattn": [u]}
    y, _ = s.sample(S, B, (4, H, W), cond, eta=1.0, x_T=x, verbose=False, unconditional_guidance_scale=2.5,
                    unconditional_conditioning=ucond)
    return y


for _ in range(2):
    one(xT, ccat, ctx, uc)
torch.cuda.synchronize()
for mode in ("no sync between calls", "sync + H2D each call"):
    t0 = time.perf_counter()
    for _ in range(3):
        if mode.startswith("sync"):
            a, b, c, d = [t.to(dev, non_blocking=True) for t in (xT_h, ccat_h, ctx_h, uc_h)]
            y = one(a, b, c, d)
            torch.cuda.synchronize()
        else:
            y = one(xT, ccat, ctx, uc)
    torch.cuda.synchronize()
    print(f"{mode}: {(time.perf_counter() - t0) / 3 * 1e3:.1f} ms per sample() call", flush=True)
# timeline of one synced call
torch.cuda.synchronize()
marks.clear()
e_in, e_out = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t_in = time.perf_counter()
e_in.record()
y = one(xT, ccat, ctx, uc)
e_out.record()
t_ret = time.perf_counter()
torch.cuda.synchronize()
t_done = time.perf_counter()
print(f"CPU: entry -> first replay {1e3 * (marks[0][0] - t_in):.2f} ms, -> last replay issued "
      f"{1e3 * (marks[S - 1][0] - t_in):.2f} ms, -> return {1e3 * (t_ret - t_in):.2f} ms, -> GPU done "
      f"{1e3 * (t_done - t_in):.2f} ms")
print(f"GPU: entry -> first replay {e_in.elapsed_time(marks[0][1]):.2f} ms, first -> last replay start "
      f"{marks[0][1].elapsed_time(marks[S - 1][1]):.2f} ms, total {e_in.elapsed_time(e_out):.2f} ms")
gaps = [marks[i][1].elapsed_time(marks[i + 1][1]) for i in range(S - 1)]
print("per-step GPU ms (first 6, min, max):", [round(g, 2) for g in gaps[:6]], round(min(gaps), 2), round(max(gaps), 2))
