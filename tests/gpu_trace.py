"""In-kernel timelines (clock64 samples of CTA 0) of the GEMM/conv and attention kernels at UNet sizes.

    LR_GEMM_TRACE=1 LR_ATTN_TRACE=1 python tests/gpu_trace.py

GEMM: per tile of CTA 0 and per warp role, cycles relative to the first sample:
  producer: start / all loads of the tile issued;  MMA: before / after accumulator-free wait, first operands landed,
  all MMAs issued;  epilogue warp 2: tile start, ready to read TMEM, accumulator full, chunks done, staging barrier,
  TMA store issued.
Attention: cycles per softmax phase, averaged per KV tile (see attention_tc.cuh).
"""
import ctypes
import os
import sys

os.environ.setdefault("LR_GEMM_TRACE", "1")
os.environ.setdefault("LR_ATTN_TRACE", "1")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch  # noqa: E402

from leftrefill_b200 import _native as N  # noqa: E402
from leftrefill_b200 import ops  # noqa: E402


def read_trace(n_u64=3 * 16 * 8):
    buf = (ctypes.c_ulonglong * n_u64)()
    N.check(N.lib().lr_debug_read_trace(ctypes.cast(buf, ctypes.c_void_p), n_u64 * 8, 1), "read_trace")
    return list(buf)


def show_gemm(name, fn, reps=3):
    for _ in range(reps):  # warm-up launches also write the trace; the last launch wins
        fn()
    torch.cuda.synchronize()
    read_trace()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    tr = read_trace()
    vals = [v for v in tr if v]
    if not vals:
        print(f"[{name}] no trace")
        return
    t0 = min(vals)
    print(f"[{name}] {e0.elapsed_time(e1) * 1e3:.1f} us; cycles since first sample, CTA 0")
    roles = ["producer", "mma", "epilogue"]
    for r in range(3):
        for t in range(16):
            row = tr[(r * 16 + t) * 8:(r * 16 + t) * 8 + 8]
            if not any(row):
                continue
            print(f"   {roles[r]:9s} tile {t:2d}: " + " ".join(f"{(v - t0) if v else -1:7d}" for v in row[:8]))


def show_attn(name, fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    read_trace()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    tr = read_trace(16)
    print(f"[{name}] {e0.elapsed_time(e1) * 1e3:.1f} us")
    names = ["wait_S", "ld_S", "max", "exp", "wait_PV", "st_P"]
    for t in range(2):
        row = tr[t * 8:t * 8 + 8]
        tiles = max(1, row[7])
        print(f"   warpgroup {t}: tiles={row[7]} total={row[6]} cyc ({row[6] / tiles:.0f}/tile)  " +
              "  ".join(f"{n}={row[i] / tiles:.0f}" for i, n in enumerate(names)))


def main():
    torch.manual_seed(0)
    a = torch.randn(65536, 320, device="cuda").half()
    w = torch.randn(320, 320, device="cuda").half() * 0.05
    r = torch.randn(65536, 320, device="cuda").half()
    b = torch.zeros(320, device="cuda")
    show_gemm("linear 65536x320->320 +bias", lambda: ops.linear(a, w, bias=b))
    show_gemm("linear 65536x320->320 +bias +res", lambda: ops.linear(a, w, bias=b, residual=r))
    show_gemm("linear 65536x320->320 +bias +res cg1 bn160", lambda: ops.linear(a, w, bias=b, residual=r,
                                                                              force_block_n=1160))
    x = torch.randn(8, 64, 128, 320, device="cuda").half()
    wt = torch.randn(320, 2880, device="cuda").half() * 0.01
    show_gemm("conv 8x64x128 320->320", lambda: ops.conv3x3(x, wt, bias=b))
    x2 = torch.randn(8, 8, 16, 1280, device="cuda").half()
    wt2 = torch.randn(1280, 11520, device="cuda").half() * 0.01
    b2 = torch.zeros(1280, device="cuda")
    show_gemm("conv 8x8x16 1280->1280", lambda: ops.conv3x3(x2, wt2, bias=b2))
    x3 = torch.randn(8, 16, 32, 1280, device="cuda").half()
    show_gemm("conv 8x16x32 1280->1280", lambda: ops.conv3x3(x3, wt2, bias=b2))
    q = torch.randn(8, 8192, 320, device="cuda").half()
    k = torch.randn(8, 8192, 320, device="cuda").half()
    v = torch.randn(8, 8192, 320, device="cuda").half()
    show_attn("attention b=8 h=5 8192x8192", lambda: ops.attention(q, k, v, 5))
    q2 = torch.randn(8, 2048, 640, device="cuda").half()
    show_attn("attention b=8 h=10 2048x2048", lambda: ops.attention(q2, q2, q2, 10))


if __name__ == "__main__":
    main()
