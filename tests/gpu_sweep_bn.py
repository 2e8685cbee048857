"""Sweep the UMMA N tile (block_n) and CTA-group for representative GEMM shapes: which N granularity is efficient?"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from leftrefill_b200 import ops  # noqa: E402


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


torch.manual_seed(0)
for (n, h, w_, c, co) in [(8, 64, 128, 320, 320), (8, 32, 64, 640, 640), (8, 16, 32, 1280, 1280)]:
    x = torch.randn(n, h, w_, c, device="cuda").half()
    wt = torch.randn(co, 9 * c, device="cuda").half() * 0.01
    b = torch.zeros(co, device="cuda")
    fl = 2.0 * n * h * w_ * 9 * c * co
    res = []
    for cg in (1, 2):
        for bn in (64, 96, 128, 160, 192, 224, 256):
            ms = timeit(lambda: ops.conv3x3(x, wt, bias=b, force_block_n=cg * 1000 + bn))
            res.append(f"cg{cg}/bn{bn}: {fl / ms / 1e9:6.0f}")
    print(f"conv {n}x{h}x{w_} {c}->{co} TFLOP/s  " + "  ".join(res), flush=True)
for (M, K, Nn) in [(65536, 320, 320), (65536, 320, 960), (16384, 640, 640), (65536, 1280, 320)]:
    a = torch.randn(M, K, device="cuda").half()
    w = torch.randn(Nn, K, device="cuda").half()
    r = torch.randn(M, Nn, device="cuda").half()
    fl = 2.0 * M * K * Nn
    res = []
    for cg in (1, 2):
        for bn in (64, 96, 128, 160, 192, 256):
            ms = timeit(lambda: ops.linear(a, w, residual=r, force_block_n=cg * 1000 + bn))
            res.append(f"cg{cg}/bn{bn}: {ms * 1e3:5.1f}us")
    print(f"linear+res {M}x{K}x{Nn}  " + "  ".join(res), flush=True)
