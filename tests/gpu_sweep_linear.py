"""(cg, block_n) sweep of every Linear shape of the UNet forward at N = 8, 64x128 (CUDA-graph timed): is the tile-cost
model of build_conv_op still right after the lean epilogue?  default = what the model picks (force_block_n = 0)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
from gpu_time_norm import timeit  # noqa: E402
from leftrefill_b200 import ops  # noqa: E402

torch.manual_seed(0)
SHAPES = [(65536, 320, 320, True, False), (65536, 320, 320, False, False), (65536, 320, 960, False, False),
          (65536, 1280, 320, True, False), (65536, 320, 1280, False, True),
          (16384, 640, 640, True, False), (16384, 640, 1920, False, False), (16384, 2560, 640, True, False),
          (16384, 640, 2560, False, True),
          (4096, 1280, 1280, True, False), (4096, 1280, 3840, False, False), (4096, 5120, 1280, True, False),
          (4096, 1280, 5120, False, True)]
for (M, K, Nn, res, geglu) in SHAPES:
    a = torch.randn(M, K, device="cuda").half()
    w = torch.randn(2 * Nn if geglu else Nn, K, device="cuda").half() * 0.03
    b = torch.zeros(2 * Nn if geglu else Nn, device="cuda")
    r = torch.randn(M, Nn, device="cuda").half() if res else None
    out = []
    best = (1e9, None)
    for force in [0] + [cg * 1000 + bn for cg in (1, 2) for bn in (96, 128, 160, 192, 224, 256)]:
        if geglu and (force % 1000) % 64 != 0:
            continue
        try:
            us = timeit([lambda: ops.linear(a, w, bias=b, residual=r, geglu=geglu, force_block_n=force)])
        except Exception as e:  # noqa: BLE001
            out.append(f"{force}: err")
            continue
        out.append(f"{'default' if force == 0 else force}: {us:5.1f}")
        if us < best[0]:
            best = (us, force)
    print(f"linear {M}x{K}->{Nn}{' +res' if res else ''}{' geglu' if geglu else ''}  best {best[1]} {best[0]:.1f} us | " + "  ".join(out),
          flush=True)
