"""Element-wise report against the tolerance BASELINE.json states (rtol 1e-3 / atol 1e-4): for each configuration the
fraction of output elements with |a - b| <= atol + rtol * |b| for
    native vs fp32 oracle | oracle under torch.autocast vs fp32 oracle | native vs oracle under autocast,
plus relative RMS / max abs. The fp32 oracle and the autocast oracle run on the GPU (same weights, same inputs).
    python tests/gpu_tolerance_table.py [out.md]
An fp16 tensor-core pipeline through ~60 layers cannot meet rtol 1e-3 element-wise end to end - the reference's own
autocast path does not (middle column) - so the tests assert "not worse than 1.25x the autocast path"; this table says
how far both are from the stated tolerance."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import O, err_stats, synthetic_inputs  # noqa: E402

import leftrefill_b200 as lr  # noqa: E402

RTOL, ATOL = 1e-3, 1e-4


def frac(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs() <= ATOL + RTOL * b.abs()).float().mean().item()


def row(name, native, auto, ref):
    sn, sa, sna = err_stats(native, ref), err_stats(auto, ref), err_stats(native, auto)
    return (f"| {name} | {frac(native, ref):.4f} | {frac(auto, ref):.4f} | {frac(native, auto):.4f} | "
            f"{sn['rel_rms']:.2e} / {sn['max_abs']:.2e} | {sa['rel_rms']:.2e} / {sa['max_abs']:.2e} | "
            f"{sna['rel_rms']:.2e} / {sna['max_abs']:.2e} |")


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else None
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda")
    lines = ["| config | native vs fp32: frac within | autocast vs fp32: frac within | native vs autocast: frac within | "
             "native vs fp32 rel-RMS / max | autocast vs fp32 rel-RMS / max | native vs autocast rel-RMS / max |",
             "|---|---|---|---|---|---|---|"]
    for name, cfg, n, h, w in [("small cfg (64 ch), 2x9x16x32", O.SMALL_CFG, 2, 16, 32),
                               ("full cfg 865.9M, 1x9x16x32", O.DEFAULT_CFG, 1, 16, 32),
                               ("full cfg, C5 size 2x9x32x64", O.DEFAULT_CFG, 2, 32, 64),
                               ("full cfg, C2 size 2x9x64x128 (one CFG pair)", O.DEFAULT_CFG, 2, 64, 128)]:
        sd = O.make_state_dict(cfg, seed=0)
        m = lr.UNetModel(**cfg)
        m.load_state_dict(sd, strict=True)
        m = m.cuda().eval()
        sdc = {k: v.cuda() for k, v in sd.items()}
        g = torch.Generator().manual_seed(99)
        x = torch.randn(n, 9, h, w, generator=g).cuda()
        ctx = torch.randn(n, 77, cfg["context_dim"], generator=g).cuda()
        t = torch.tensor([981, 401][:n]).cuda()
        with torch.no_grad():
            ref = O.unet_forward(sdc, cfg, x, t, ctx)
            with torch.autocast("cuda"):
                auto = O.unet_forward(sdc, cfg, x, t, ctx).float()
            y = m(x, t, context=ctx)
        lines.append(row(name, y, auto, ref))
        print(lines[-1], flush=True)
        del m, sdc, sd
        torch.cuda.empty_cache()
    text = (f"Element-wise tolerance report, rtol {RTOL} / atol {ATOL} (BASELINE.json north_star), UNet forward eps "
            f"output, B200.\n\n" + "\n".join(lines) + "\n")
    print(text)
    if out_path:
        open(out_path, "w").write(text)


if __name__ == "__main__":
    main()
