"""Where the persistent GroupNorm's time goes: LR_GN_DEBUG=1 skips pass 1 (statistics), 2 skips pass 2 (apply), 3 both
(launch + barrier only). Each setting runs in a subprocess (the switch is read once per process)."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def run():
    import torch
    from gpu_time_norm import timeit
    from leftrefill_b200 import ops
    tag = os.environ.get("LR_GN_DEBUG", "0")
    shapes = [(8, 64, 128, 320, 0, True), (8, 64, 128, 320, 0, False), (8, 32, 64, 640, 0, True),
              (8, 64, 128, 640, 320, True), (2, 64, 128, 320, 0, True)]
    if os.environ.get("LR_GN_SHAPES"):
        shapes = shapes[:int(os.environ["LR_GN_SHAPES"])]
    for (n, h, w, c0, c1, silu) in shapes:
        C = c0 + c1
        nbytes = n * h * w * C * 2
        nbuf = max(2, int(200e6 // nbytes) + 1)
        xs0 = [torch.randn(n, h, w, c0, device="cuda").half() for _ in range(nbuf)]
        xs1 = [torch.randn(n, h, w, c1, device="cuda").half() for _ in range(nbuf)] if c1 else [None] * nbuf
        g = torch.randn(C, device="cuda")
        b = torch.randn(C, device="cuda")
        hot = timeit([lambda: ops.groupnorm(xs0[0], g, b, 1e-5, silu=silu, x1=xs1[0])])
        cold = timeit([(lambda i=i: ops.groupnorm(xs0[i], g, b, 1e-5, silu=silu, x1=xs1[i])) for i in range(nbuf)])
        print(f"dbg={tag} GN n={n} {h}x{w} c={c0}+{c1} silu={int(silu)}: hot {hot:6.1f} us  cold {cold:6.1f} us", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run()
    else:
        for dbg in os.environ.get("LR_GN_PASSES", "0 1 2 3").split():
            e = dict(os.environ, LR_GN_DEBUG=dbg, LR_GN_FUSED_KB="0")
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "run"], env=e, capture_output=True, text=True,
                               timeout=300)
            print(r.stdout.strip() or r.stderr.strip()[-800:], flush=True)
