"""Where do the ~600 cycles per k-chunk go? LR_GEMM_DEBUG bits: 1 = no A loads, 2 = no B loads, 4 = no MMAs."""
import os
import subprocess
import sys

if len(sys.argv) > 1:
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from leftrefill_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(8, 64, 128, 320, device="cuda").half()
    wt = torch.randn(320, 2880, device="cuda").half() * 0.01
    b = torch.zeros(320, device="cuda")
    out = []
    for f in (1160, 1256, 2160, 2256, 1064):
        fn = lambda: ops.conv3x3(x, wt, bias=b, force_block_n=f)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(f"{f}: {e0.elapsed_time(e1) * 100:6.1f}us")
    print(f"dbg={os.environ.get('LR_GEMM_DEBUG', '0')}  conv320: " + "  ".join(out), flush=True)
    a = torch.randn(65536, 320, device="cuda").half()
    w = torch.randn(320, 320, device="cuda").half() * 0.05
    r = torch.randn(65536, 320, device="cuda").half()
    out = []
    for f, res in ((1160, None), (2160, None), (1160, r), (2160, r), (1064, None), (1256, None)):
        fn = lambda: ops.linear(a, w, bias=b, residual=res, force_block_n=f)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(f"{f}{'+res' if res is not None else ''}: {e0.elapsed_time(e1) * 100:6.1f}us")
    print(f"dbg={os.environ.get('LR_GEMM_DEBUG', '0')}  lin320: " + "  ".join(out), flush=True)
else:
    for dbg in [int(v) for v in os.environ.get('LR_DBG_LIST', '0,3,4,7').split(',')]:
        env = dict(os.environ, LR_GEMM_DEBUG=str(dbg))
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "run"], env=env, capture_output=True, text=True,
                           timeout=120)
        print((r.stdout.strip() or r.stderr.strip()[-300:]), flush=True)
