"""In-kernel timelines (LR_GEMM_TRACE) of the small-K Linears of the 64x128 level: where one tile period goes.

    python tests/gpu_trace_lin.py            (LR_NO_LEAN_EPI=1 for the general epilogue path)
"""
import os
import sys

os.environ.setdefault("LR_GEMM_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch  # noqa: E402

from gpu_trace import show_gemm  # noqa: E402
from leftrefill_b200 import ops  # noqa: E402


def main():
    torch.manual_seed(0)
    a = torch.randn(65536, 320, device="cuda").half()
    w = torch.randn(320, 320, device="cuda").half() * 0.05
    r = torch.randn(65536, 320, device="cuda").half()
    b = torch.zeros(320, device="cuda")
    show_gemm("linear 65536x320->320 +bias", lambda: ops.linear(a, w, bias=b))
    show_gemm("linear 65536x320->320 +bias +res", lambda: ops.linear(a, w, bias=b, residual=r))
    wq = torch.randn(960, 320, device="cuda").half() * 0.05
    show_gemm("linear 65536x320->960 (qkv, no bias)", lambda: ops.linear(a, wq))
    wg = torch.randn(2560, 320, device="cuda").half() * 0.05
    bg = torch.zeros(2560, device="cuda")
    show_gemm("linear 65536x320->2560 geglu", lambda: ops.linear(a, wg, bias=bg, geglu=True))
    a4 = torch.randn(65536, 1280, device="cuda").half()
    w4 = torch.randn(320, 1280, device="cuda").half() * 0.03
    show_gemm("linear 65536x1280->320 +bias +res", lambda: ops.linear(a4, w4, bias=b, residual=r))
    a2 = torch.randn(16384, 640, device="cuda").half()
    w2 = torch.randn(640, 640, device="cuda").half() * 0.04
    r2 = torch.randn(16384, 640, device="cuda").half()
    b2 = torch.zeros(640, device="cuda")
    show_gemm("linear 16384x640->640 +bias +res", lambda: ops.linear(a2, w2, bias=b2, residual=r2))


if __name__ == "__main__":
    main()
