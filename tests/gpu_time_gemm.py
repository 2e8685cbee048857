"""Kernel-only timings of gemm_conv_kernel / attention at the UNet's shapes (CUDA-graph replay of 20 launches, so the
Python wrapper cost is not in the number).   python tests/gpu_time_gemm.py [tag]
Set LR_B200_LIB=<path to another liblr_b200.so> to time a different build of the same ABI."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from leftrefill_b200 import ops  # noqa: E402

TAG = sys.argv[1] if len(sys.argv) > 1 else "build"


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    torch.manual_seed(0)
    rows = []
    for (M, K, Nn, res, geglu) in [(65536, 320, 320, False, False), (65536, 320, 320, True, False),
                                   (65536, 320, 960, False, False), (65536, 320, 1280, False, True),
                                   (65536, 1280, 320, True, False), (16384, 640, 640, True, False),
                                   (16384, 640, 1920, False, False), (16384, 640, 2560, False, True),
                                   (4096, 1280, 1280, True, False), (4096, 1280, 5120, False, True)]:
        a = torch.randn(M, K, device="cuda").half()
        w = torch.randn(2 * Nn if geglu else Nn, K, device="cuda").half() * 0.05
        b = torch.zeros(2 * Nn if geglu else Nn, device="cuda")
        r = torch.randn(M, Nn, device="cuda").half() if res else None
        us = timeit(lambda: ops.linear(a, w, bias=b, residual=r, geglu=geglu))
        fl = 2.0 * M * K * (2 * Nn if geglu else Nn)
        rows.append((f"linear {M}x{K}->{Nn}{' +res' if res else ''}{' geglu' if geglu else ''}", us, fl))
    for (n, h, w_, c, co, res) in [(8, 64, 128, 320, 320, True), (8, 32, 64, 640, 640, True),
                                   (8, 16, 32, 1280, 1280, True), (8, 8, 16, 1280, 1280, True),
                                   (8, 64, 128, 640, 320, False), (8, 32, 64, 1280, 1280, False)]:
        x = torch.randn(n, h, w_, c, device="cuda").half()
        wt = torch.randn(co, 9 * c, device="cuda").half() * 0.01
        b = torch.zeros(co, device="cuda")
        r = torch.randn(n, h, w_, co, device="cuda").half() if res else None
        us = timeit(lambda: ops.conv3x3(x, wt, bias=b, residual=r))
        rows.append((f"conv {n}x{h}x{w_} {c}->{co}{' +res' if res else ''}", us, 2.0 * n * h * w_ * 9 * c * co))
    for (b_, hd, t, tk) in [(8, 5, 8192, 8192), (8, 10, 2048, 2048), (8, 20, 512, 512), (8, 5, 8192, 77)]:
        q = torch.randn(b_, t, hd * 64, device="cuda").half()
        k = torch.randn(b_, tk, hd * 64, device="cuda").half()
        v = torch.randn(b_, tk, hd * 64, device="cuda").half()
        us = timeit(lambda: ops.attention(q, k, v, hd), iters=5)
        rows.append((f"attention b={b_} h={hd} {t}x{tk}", us, 4.0 * b_ * hd * t * tk * 64))
    for name, us, fl in rows:
        print(f"{TAG:>10s} {name:42s} {us:8.1f} us {fl / us / 1e6:8.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
