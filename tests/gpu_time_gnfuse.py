"""A/B timing of the fused-GroupNorm GEMM paths at UNet sizes (op level): plain conv / linear vs the same op with producer
statistics, with the in-smem transform, and with the timing-experiment bits of LR_GEMM_DEBUG (64: transform warps do the
barrier hand-offs only; 128: cluster-scope release / acquire on the hand-off barrier).
    python tests/gpu_time_gnfuse.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from leftrefill_b200 import ops  # noqa: E402


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    torch.manual_seed(0)
    for (n, h, w, c, co) in [(8, 64, 128, 320, 320), (8, 32, 64, 640, 640), (8, 16, 32, 1280, 1280), (8, 64, 128, 640, 320)]:
        x = torch.randn(n, h, w, c, device="cuda").half()
        wt = (torch.randn(co, 9 * c, device="cuda") * 0.01).half()
        b = torch.zeros(co, device="cuda")
        r = torch.randn(n, h, w, co, device="cuda").half()
        sc = torch.rand(n, c, device="cuda") + 0.5
        sh = torch.randn(n, c, device="cuda") * 0.1
        fl = 2.0 * n * h * w * 9 * c * co
        rows = {}
        os.environ.pop("LR_GEMM_DEBUG", None)
        rows["plain"] = timeit(lambda: ops.conv3x3(x, wt, bias=b, residual=r))
        rows["stats"] = timeit(lambda: ops.gn_conv3x3(x, wt, bias=b, residual=r, want_stats=True))
        for tag, dbg in [("xf", None), ("xf nomath(64)", "64"), ("xf cluster(128)", "128"), ("xf nomath+cluster(192)", "192")]:
            if dbg is None:
                os.environ.pop("LR_GEMM_DEBUG", None)
            else:
                os.environ["LR_GEMM_DEBUG"] = dbg
            rows[tag] = timeit(lambda: ops.gn_conv3x3(x, wt, gn=(sc, sh), silu=True, bias=b, residual=r))
            rows[tag + " nosilu"] = timeit(lambda: ops.gn_conv3x3(x, wt, gn=(sc, sh), silu=False, bias=b, residual=r))
        os.environ.pop("LR_GEMM_DEBUG", None)
        rows["xf+stats"] = timeit(lambda: ops.gn_conv3x3(x, wt, gn=(sc, sh), silu=True, bias=b, residual=r, want_stats=True))
        print(f"conv3x3 n={n} {h}x{w} {c}->{co} (+res): " +
              " | ".join(f"{k} {v:.1f} us ({fl / v / 1e6:.0f} TF)" for k, v in rows.items()), flush=True)
    for (M, P, K, Nn) in [(65536, 8192, 320, 320), (16384, 2048, 640, 640), (4096, 512, 1280, 1280)]:
        a = torch.randn(M, K, device="cuda").half()
        wt = (torch.randn(Nn, K, device="cuda") * 0.02).half()
        b = torch.zeros(Nn, device="cuda")
        r = torch.randn(M, Nn, device="cuda").half()
        sc = torch.rand(M // P, K, device="cuda") + 0.5
        sh = torch.randn(M // P, K, device="cuda") * 0.1
        fl = 2.0 * M * K * Nn
        rows = {}
        os.environ.pop("LR_GEMM_DEBUG", None)
        rows["plain"] = timeit(lambda: ops.linear(a, wt, bias=b))
        rows["plain+res"] = timeit(lambda: ops.linear(a, wt, bias=b, residual=r))
        rows["+res stats"] = timeit(lambda: ops.gn_linear(a, wt, P, bias=b, residual=r, want_stats=True))
        for tag, dbg in [("xf", None), ("xf nomath(64)", "64"), ("xf cluster(128)", "128")]:
            if dbg is None:
                os.environ.pop("LR_GEMM_DEBUG", None)
            else:
                os.environ["LR_GEMM_DEBUG"] = dbg
            rows[tag] = timeit(lambda: ops.gn_linear(a, wt, P, gn=(sc, sh), silu=False, bias=b))
        os.environ.pop("LR_GEMM_DEBUG", None)
        print(f"linear {M}x{K}->{Nn}: " + " | ".join(f"{k} {v:.1f} us ({fl / v / 1e6:.0f} TF)" for k, v in rows.items()),
              flush=True)


if __name__ == "__main__":
    main()
