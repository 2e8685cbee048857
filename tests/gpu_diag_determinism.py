"""Run-to-run determinism and batch-position invariance of each op and of the UNet (bring-up diagnostics)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import O, err_stats, synthetic_inputs  # noqa: E402

import leftrefill_b200 as lr  # noqa: E402
from leftrefill_b200 import ops  # noqa: E402


def same(tag, a, b):
    d = (a.float() - b.float()).abs()
    nz = (d > 0).float().mean().item()
    print(f"{tag}: max|diff|={d.max().item():.3e} frac_diff={nz:.5f} rel_rms={(d.pow(2).mean().sqrt() / a.float().pow(2).mean().sqrt()).item():.3e}",
          flush=True)


def main():
    torch.manual_seed(0)
    dev = "cuda"
    # ---- ops ----
    a = torch.randn(8192, 320, device=dev).half()
    w = torch.randn(320, 320, device=dev).half() * 0.05
    r = torch.randn(8192, 320, device=dev).half()
    b = torch.randn(320, device=dev)
    same("linear+res twice", ops.linear(a, w, bias=b, residual=r), ops.linear(a, w, bias=b, residual=r))
    w2 = torch.randn(2560, 320, device=dev).half() * 0.05
    same("geglu twice", ops.linear(a, w2, geglu=True), ops.linear(a, w2, geglu=True))
    x = torch.randn(8, 32, 64, 320, device=dev).half()
    wc = torch.randn(320, 9 * 320, device=dev).half() * 0.02
    same("conv twice", ops.conv3x3(x, wc, bias=b), ops.conv3x3(x, wc, bias=b))
    perm = torch.tensor([5, 2, 7, 0, 3, 6, 1, 4], device=dev)
    same("conv batch-perm", ops.conv3x3(x[perm].contiguous(), wc, bias=b), ops.conv3x3(x, wc, bias=b)[perm])
    g, be = torch.randn(320, device=dev), torch.randn(320, device=dev)
    same("groupnorm twice", ops.groupnorm(x, g, be, 1e-5, silu=True), ops.groupnorm(x, g, be, 1e-5, silu=True))
    same("groupnorm batch-perm", ops.groupnorm(x[perm].contiguous(), g, be, 1e-5, silu=True),
         ops.groupnorm(x, g, be, 1e-5, silu=True)[perm])
    same("layernorm twice", ops.layernorm(a, g, be), ops.layernorm(a, g, be))
    q = torch.randn(8, 2048, 640, device=dev).half()
    k = torch.randn(8, 2048, 640, device=dev).half()
    v = torch.randn(8, 2048, 640, device=dev).half()
    o1 = ops.attention(q, k, v, 10)
    same("attention twice", o1, ops.attention(q, k, v, 10))
    same("attention batch-perm", ops.attention(q[perm].contiguous(), k[perm].contiguous(), v[perm].contiguous(), 10),
         o1[perm])
    kc = torch.randn(8, 77, 640, device=dev).half()
    vc = torch.randn(8, 77, 640, device=dev).half()
    same("cross-attention twice", ops.attention(q, kc, vc, 10), ops.attention(q, kc, vc, 10))
    q1 = torch.randn(8, 128, 1280, device=dev).half()
    same("attention T=128 twice", ops.attention(q1, q1, q1, 20), ops.attention(q1, q1, q1, 20))

    # ---- UNet small ----
    for cfg, hw, nb in [(O.SMALL_CFG, (16, 32), 4), (O.DEFAULT_CFG, (64, 128), 4)]:
        m = lr.UNetModel(**cfg)
        m.load_state_dict(O.make_state_dict(cfg, seed=0), strict=True)
        m = m.cuda().eval()
        xT, c_cat, ctx, uc = synthetic_inputs(nb, h=hw[0], w=hw[1], ctx_dim=cfg["context_dim"], device=dev)
        xc = torch.cat([torch.cat([xT, xT]), torch.cat([c_cat, c_cat])], dim=1).contiguous()
        cc = torch.cat([uc, ctx]).contiguous()
        t = torch.full((2 * nb,), 981, dtype=torch.long, device=dev)
        with torch.no_grad():
            y1 = m(xc, t, context=cc)
            y2 = m(xc, t, context=cc)
            same(f"UNet mc={cfg['model_channels']} twice", y1, y2)
            p = torch.randperm(2 * nb, device=dev)
            yp = m(xc[p].contiguous(), t, context=cc[p].contiguous())
            same(f"UNet mc={cfg['model_channels']} batch-perm", yp, y1[p])
            y3 = m(xc, t, context=torch.cat([ctx, ctx]).contiguous())
            same(f"UNet mc={cfg['model_channels']} cond==uncond halves", y3[:nb], y3[nb:])
        del m
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
