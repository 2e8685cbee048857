"""GPU bring-up diagnostics for the whole UNet / sampler path (prints numbers; the pytest files assert on them).

    python tests/gpu_diag_unet.py [--skip-full]
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import FakeLDM, O, err_stats, load_golden, synthetic_inputs  # noqa: E402

from leftrefill_b200 import _native as N  # noqa: E402
from leftrefill_b200.ddim import DDIMSampler  # noqa: E402
from leftrefill_b200.unet import MultiViewUnetModel, UNetModel  # noqa: E402


def build(cfg, seed, cls=UNetModel, **kw):
    sd = O.make_state_dict(cfg, seed=seed)
    m = cls(**cfg, **kw)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval(), sd


def autocast_floor(sd, cfg, x, t, ctx, **kw):
    """The oracle run the way the reference runs on a GPU: torch.autocast fp16 over fp32 weights."""
    sdc = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad(), torch.autocast("cuda"):
        return O.unet_forward(sdc, cfg, x.cuda(), t.cuda(), ctx.cuda(), **kw).float()


def show(tag, got, ref):
    s = err_stats(got, ref)
    print(f"{tag}: max_abs={s['max_abs']:.3e} rms={s['rms']:.3e} ref_max={s['ref_max']:.3e} ref_rms={s['ref_rms']:.3e} "
          f"rel_rms={s['rel_rms']:.3e} finite={s['finite']}", flush=True)
    return s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-full", action="store_true")
    args = ap.parse_args()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = "cuda"

    # ---- small config vs golden fixture from the reference ----
    cfg = O.SMALL_CFG
    m, sd = build(cfg, 0)
    for name in ["unet_small.npz", "unet_small_ragged.npz"]:
        g = load_golden(name)
        x, t, ctx = torch.tensor(g["x"]), torch.tensor(g["t"]), torch.tensor(g["context"])
        with torch.no_grad():
            y = m(x.to(dev), t.to(dev), context=ctx.to(dev))
        torch.cuda.synchronize()
        show(f"[{name}] native vs reference-fp32 golden", y, g["out"])
        show(f"[{name}] torch-autocast(oracle) vs golden  ", autocast_floor(sd, cfg, x, t, ctx), g["out"])
    print("engine flops (small, last plan):", m.last_flops())

    # ---- multiview ----
    for name in ["multiview_v2.npz"]:
        g = load_golden(name)
        mv, sdm = build(cfg, 1, MultiViewUnetModel, view_num=int(g["view_num"]), concat_target=bool(g["concat_target"]))
        x, t, ctx = torch.tensor(g["x"]), torch.tensor(g["t"]), torch.tensor(g["context"])
        with torch.no_grad():
            y = mv(x.to(dev), t.to(dev), context=ctx.to(dev))
        show(f"[{name}] native vs golden", y, g["out"])
        show(f"[{name}] autocast floor  ", autocast_floor(sdm, cfg, x, t, ctx, view_num=int(g["view_num"]),
                                                          concat_target=bool(g["concat_target"])), g["out"])
        del mv

    # ---- DDIM sampler vs golden ----
    g = load_golden("ddim_small.npz")
    ldm = FakeLDM(m, torch.device(dev))
    for tag, eta in [("eta0", 0.0), ("eta1", 1.0)]:
        noises = torch.tensor(g[f"{tag}.noises"]).to(dev)
        s = DDIMSampler(ldm)
        s.noise_source = lambda shape, device, i: noises[i]
        cond = {"c_concat": [torch.tensor(g["c_concat"]).to(dev)], "c_crossattn": [torch.tensor(g["context"]).to(dev)]}
        uc = {"c_concat": [torch.tensor(g["c_concat"]).to(dev)], "c_crossattn": [torch.tensor(g["uc_context"]).to(dev)]}
        samples, inter = s.sample(4, 1, (4, 16, 32), cond, eta=eta, x_T=torch.tensor(g["x_T"]).to(dev), verbose=False,
                                  unconditional_guidance_scale=2.5, unconditional_conditioning=uc, log_every_t=1)
        show(f"[ddim {tag}] native sampler vs reference golden", samples, g[f"{tag}.samples"])
        print(f"   n_inter={len(inter['x_inter'])} (golden {int(g[f'{tag}.n_inter'])})")
    del m
    torch.cuda.empty_cache()
    if args.skip_full:
        return

    # ---- full config ----
    cfg = O.DEFAULT_CFG
    t0 = time.time()
    m, sd = build(cfg, 0)
    print(f"full model built in {time.time() - t0:.1f}s")
    g = load_golden("unet_full_16x32.npz")
    x, t, ctx = torch.tensor(g["x"]), torch.tensor(g["t"]), torch.tensor(g["context"])
    with torch.no_grad():
        y = m(x.to(dev), t.to(dev), context=ctx.to(dev))
    show("[unet_full_16x32] native vs reference golden", y, g["out"])
    show("[unet_full_16x32] autocast floor            ", autocast_floor(sd, cfg, x, t, ctx), g["out"])

    # ---- full size timing: N = 8 (4 canvases + CFG), 64x128 ----
    xT, c_cat, ctx, uc = synthetic_inputs(4, device=dev)
    xc = torch.cat([torch.cat([xT, xT]), torch.cat([c_cat, c_cat])], dim=1).contiguous()
    cc = torch.cat([uc, ctx]).contiguous()
    tt = torch.full((8,), 981, dtype=torch.long, device=dev)
    m.sync_weights()
    m.set_context(cc)
    for _ in range(3):
        y = m.forward_native(xc, tt, None)
    torch.cuda.synchronize()
    N.lib().lr_launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    iters = 10
    for _ in range(iters):
        y = m.forward_native(xc, tt, None)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = m.last_flops()
    print(f"[full N=8 64x128] {ms:.2f} ms/step, {fl / 1e12:.3f} TFLOP -> {fl / ms / 1e9:.1f} TFLOP/s, "
          f"launches/step={N.lib().lr_launch_count() / iters:.0f}, device MB={N.lib().lr_unet_device_bytes(m.engine()) / 2**20:.0f}",
          flush=True)
    print("finite:", bool(torch.isfinite(y).all()), "out rms", y.pow(2).mean().sqrt().item())
    # parity at full size against the oracle on the GPU in fp32 (B=2 to bound memory: fp32 logits 2x5x8192^2x4 B)
    sdc = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        y_or = O.unet_forward(sdc, cfg, xc[[0, 4]], tt[:2], cc[[0, 4]])
        y_nat = m(xc[[0, 4]], tt[:2], context=cc[[0, 4]])
        with torch.autocast("cuda"):
            t1 = time.time()
            y_ac = O.unet_forward(sdc, cfg, xc[[0, 4]], tt[:2], cc[[0, 4]]).float()
            torch.cuda.synchronize()
            print(f"torch autocast oracle B=2: {time.time() - t1:.3f}s (first call)")
    show("[full 64x128 B=2] native vs fp32 oracle(GPU)", y_nat, y_or)
    show("[full 64x128 B=2] autocast floor            ", y_ac, y_or)
    # GPU reference timing (torch ops under autocast, B=2 and B=8 if memory allows)
    for B, xs, ts, cs in [(2, xc[[0, 4]], tt[:2], cc[[0, 4]]), (8, xc, tt, cc)]:
        try:
            with torch.no_grad(), torch.autocast("cuda"):
                for _ in range(2):
                    O.unet_forward(sdc, cfg, xs, ts, cs)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(3):
                    O.unet_forward(sdc, cfg, xs, ts, cs)
                e1.record()
                torch.cuda.synchronize()
            print(f"[GPU torch-autocast oracle B={B}] {e0.elapsed_time(e1) / 3:.1f} ms/step")
        except Exception as ex:  # noqa: BLE001
            print(f"[GPU torch-autocast oracle B={B}] failed: {type(ex).__name__}")
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
