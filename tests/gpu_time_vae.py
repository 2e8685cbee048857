"""Timing of the native first-stage decoder at the bench workload (4 latents 64x128 -> 4 images 512x1024).
    python tests/gpu_time_vae.py [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import vae_oracle as V  # noqa: E402  (test tooling: synthetic weights only)

import leftrefill_b200 as lr  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cfg = V.DEFAULT_CFG
m = lr.AutoencoderKL(ddconfig={k: v for k, v in cfg.items() if k != "embed_dim"}, embed_dim=cfg["embed_dim"])
m.load_state_dict(V.make_state_dict(cfg, seed=0), strict=True)
m = m.cuda().eval()
z = torch.randn(B, 4, 64, 128, device="cuda") * 0.7
for _ in range(3):
    y = m.decode(z, z_scale=1.0 / V.SCALE_FACTOR)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    y = m.decode(z, z_scale=1.0 / V.SCALE_FACTOR)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
fl = m.last_flops()
from leftrefill_b200 import _native as N  # noqa: E402
print(f"vae decode B={B}: {ms:.2f} ms, {fl / 1e12:.2f} TFLOP -> {fl / ms / 1e9:.0f} TFLOP/s, "
      f"{N.lib().lr_vae_num_steps(m.engine())} plan steps, "
      f"{N.lib().lr_vae_device_bytes(m.engine()) / 2 ** 30:.2f} GiB device memory", flush=True)
sd = {k: p.detach() for k, p in m.named_parameters()}
with torch.no_grad(), torch.autocast("cuda"):
    for _ in range(2):
        V.decode(sd, cfg, z, scale_factor=V.SCALE_FACTOR)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        V.decode(sd, cfg, z, scale_factor=V.SCALE_FACTOR)
    e1.record()
    torch.cuda.synchronize()
print(f"context: oracle (PyTorch ops, autocast) decode B={B}: {e0.elapsed_time(e1) / 3:.2f} ms", flush=True)
