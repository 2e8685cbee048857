"""LR_GN_DEBUG=16 on a build with -DLR_GN_TRACE=1 (LR_B200_LIB=...): the persistent GroupNorm prints clock64 deltas of its
phases (CTA 0 and CTA 77). The shipped build carries no trace code (it costs registers)."""
import os, sys
os.environ.setdefault("LR_GN_DEBUG", "16")
os.environ.setdefault("LR_GN_FUSED_KB", "0")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from leftrefill_b200 import ops
for (n, h, w, c, silu) in [(8, 64, 128, 320, True), (8, 64, 128, 320, False), (8, 32, 64, 640, True)]:
    x = torch.randn(n, h, w, c, device="cuda").half()
    g = torch.randn(c, device="cuda"); b = torch.randn(c, device="cuda")
    print(f"--- n={n} {h}x{w} c={c} silu={silu}", flush=True)
    for _ in range(3):
        ops.groupnorm(x, g, b, 1e-5, silu=silu)
        torch.cuda.synchronize()
