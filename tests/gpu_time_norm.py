"""GroupNorm / LayerNorm timings at the UNet's call-site shapes (N = 8), per scheme.

    python tests/gpu_time_norm.py                # runs every scheme in a subprocess (env vars are read once per process)

Each shape is timed twice: 'hot' = same buffer every iteration (the activation was just written by the producing
kernel: L2 resident when it fits) and 'cold' = rotating over buffers totalling > 126 MB (L2 misses).
GB/s = algorithmic bytes (read once + write once) / time.
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

GN_SHAPES = [(8, 64, 128, 320, 0), (8, 64, 128, 640, 320), (8, 32, 64, 640, 0), (8, 32, 64, 1280, 640),
             (8, 16, 32, 1280, 0), (8, 16, 32, 1280, 1280), (8, 8, 16, 1280, 0), (8, 8, 16, 1280, 1280)]
LN_SHAPES = [(65536, 320), (16384, 640), (4096, 1280), (1024, 1280)] if os.environ.get("LR_TAG", "default") == "default" else []


def timeit(fns, iters=20):
    """The op wrappers allocate their outputs (CPU cost ~ the kernel time at these sizes), so the launches are
    captured in a CUDA graph and the replay is timed."""
    import torch
    for f in fns:
        f()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(graph):
            for i in range(iters):
                fns[i % len(fns)]()
    except Exception as e:  # noqa: BLE001
        print("graph capture failed, timing eagerly:", str(e)[:100])
        graph = None
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if graph is not None:
        graph.replay()
        torch.cuda.synchronize()
    e0.record()
    if graph is not None:
        graph.replay()
    else:
        for i in range(iters):
            fns[i % len(fns)]()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def run():
    import torch
    from leftrefill_b200 import ops
    torch.manual_seed(0)
    tag = os.environ.get("LR_TAG", "default")
    for (n, h, w, c0, c1) in GN_SHAPES:
        C = c0 + c1
        nbytes = n * h * w * C * 2
        nbuf = max(2, int(200e6 // nbytes) + 1)
        xs0 = [torch.randn(n, h, w, c0, device="cuda").half() for _ in range(nbuf)]
        xs1 = [torch.randn(n, h, w, c1, device="cuda").half() for _ in range(nbuf)] if c1 else [None] * nbuf
        g = torch.randn(C, device="cuda")
        b = torch.randn(C, device="cuda")
        hot = timeit([lambda: ops.groupnorm(xs0[0], g, b, 1e-5, silu=True, x1=xs1[0])])
        cold = timeit([(lambda i=i: ops.groupnorm(xs0[i], g, b, 1e-5, silu=True, x1=xs1[i])) for i in range(nbuf)])
        print(f"{tag:>14s} GN n={n} {h}x{w} c={c0}+{c1}: hot {hot:6.1f} us ({2 * nbytes / hot / 1e3:6.0f} GB/s)  "
              f"cold {cold:6.1f} us ({2 * nbytes / cold / 1e3:6.0f} GB/s)", flush=True)
        del xs0, xs1
    for (M, C) in LN_SHAPES:
        nbytes = M * C * 2
        nbuf = max(2, int(200e6 // nbytes) + 1)
        xs = [torch.randn(M, C, device="cuda").half() for _ in range(nbuf)]
        g = torch.randn(C, device="cuda")
        b = torch.randn(C, device="cuda")
        hot = timeit([lambda: ops.layernorm(xs[0], g, b)])
        cold = timeit([(lambda i=i: ops.layernorm(xs[i], g, b)) for i in range(nbuf)])
        print(f"{tag:>14s} LN M={M} C={C}: hot {hot:6.1f} us ({2 * nbytes / hot / 1e3:6.0f} GB/s)  "
              f"cold {cold:6.1f} us ({2 * nbytes / cold / 1e3:6.0f} GB/s)", flush=True)
        del xs


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run()
    else:
        schemes = [("default", {}), ("cluster<=3.5MB", {"LR_GN_FUSED_KB": "448"}), ("persistent", {"LR_GN_FUSED_KB": "0"})]
        for tag, env in schemes:
            e = dict(os.environ, LR_TAG=tag, **env)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "run"], env=e, capture_output=True,
                               text=True, timeout=300)
            print(r.stdout.strip() or r.stderr.strip()[-500:], flush=True)
