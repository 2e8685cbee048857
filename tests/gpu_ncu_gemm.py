"""Launches of the small-K GEMM shapes for an `ncu --set full --import-source on` capture.
    ncu --set full --clock-control none --import-source on -k regex:gemm_conv -s 4 -c 4 -o gpurun_out/gemm python tests/gpu_ncu_gemm.py"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from leftrefill_b200 import ops  # noqa: E402

torch.manual_seed(0)
a = torch.randn(65536, 320, device="cuda").half()
w = torch.randn(320, 320, device="cuda").half() * 0.05
wg = torch.randn(2560, 320, device="cuda").half() * 0.05
r = torch.randn(65536, 320, device="cuda").half()
b = torch.zeros(320, device="cuda")
bg = torch.zeros(2560, device="cuda")
x = torch.randn(8, 64, 128, 320, device="cuda").half()
wt = torch.randn(320, 2880, device="cuda").half() * 0.01
for _ in range(2):  # launches 0-3 warm up, 4-7 are captured
    ops.linear(a, w, bias=b)
    ops.linear(a, w, bias=b, residual=r)
    ops.linear(a, wg, bias=bg, geglu=True)
    ops.conv3x3(x, wt, bias=b, residual=r.view(8, 64, 128, 320))
torch.cuda.synchronize()
