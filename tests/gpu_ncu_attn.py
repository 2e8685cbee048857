"""Single-kernel driver for ncu captures: python tests/gpu_ncu_attn.py [attn|conv|geglu|lin320|gn|gnconv|gnlin]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from leftrefill_b200 import ops  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "attn"
torch.manual_seed(0)
if kind == "attn":
    qkv = torch.randn(8, 8192, 960, device="cuda").half()
    q, k, v = qkv[:, :, :320], qkv[:, :, 320:640], qkv[:, :, 640:]
    for _ in range(3):
        ops.attention(q, k, v, 5)
elif kind == "conv":
    x = torch.randn(8, 64, 128, 320, device="cuda").half()
    w = torch.randn(320, 2880, device="cuda").half() * 0.02
    b = torch.zeros(320, device="cuda")
    for _ in range(3):
        ops.conv3x3(x, w, bias=b)
elif kind == "gnconv":      # fused GroupNorm + SiLU conv (transform warps), preceded by the plain conv for comparison
    x = torch.randn(8, 64, 128, 320, device="cuda").half()
    w = torch.randn(320, 2880, device="cuda").half() * 0.02
    b = torch.zeros(320, device="cuda")
    r = torch.randn(8, 64, 128, 320, device="cuda").half()
    sc = torch.rand(8, 320, device="cuda") + 0.5
    sh = torch.randn(8, 320, device="cuda") * 0.1
    for _ in range(3):
        ops.conv3x3(x, w, bias=b, residual=r)
    for _ in range(3):
        ops.gn_conv3x3(x, w, gn=(sc, sh), silu=True, bias=b, residual=r, want_stats=True)
elif kind == "gn":          # persistent GroupNorm + SiLU, 8 x 64x128 x 320
    x = torch.randn(8, 64, 128, 320, device="cuda").half()
    g = torch.randn(320, device="cuda")
    b = torch.randn(320, device="cuda")
    for _ in range(3):
        ops.groupnorm(x, g, b, 1e-5, silu=True)
elif kind == "gnlin":
    a = torch.randn(65536, 320, device="cuda").half()
    w = torch.randn(320, 320, device="cuda").half() * 0.05
    b = torch.zeros(320, device="cuda")
    sc = torch.rand(8, 320, device="cuda") + 0.5
    sh = torch.randn(8, 320, device="cuda") * 0.1
    for _ in range(3):
        ops.linear(a, w, bias=b)
    for _ in range(3):
        ops.gn_linear(a, w, 8192, gn=(sc, sh), silu=False, bias=b)
elif kind == "conv1280":
    x = torch.randn(8, 32, 64, 1280, device="cuda").half()
    w = torch.randn(1280, 11520, device="cuda").half() * 0.01
    b = torch.zeros(1280, device="cuda")
    for _ in range(3):
        ops.conv3x3(x, w, bias=b)
elif kind == "geglu":
    a = torch.randn(65536, 320, device="cuda").half()
    w = torch.randn(2560, 320, device="cuda").half() * 0.05
    b = torch.zeros(2560, device="cuda")
    for _ in range(3):
        ops.linear(a, w, bias=b, geglu=True)
elif kind == "lin320":
    a = torch.randn(65536, 320, device="cuda").half()
    w = torch.randn(320, 320, device="cuda").half() * 0.05
    r = torch.randn(65536, 320, device="cuda").half()
    b = torch.zeros(320, device="cuda")
    for _ in range(3):
        ops.linear(a, w, bias=b, residual=r)
torch.cuda.synchronize()
