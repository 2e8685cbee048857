"""ORACLE TOOLING — runs ONLY in the build container (needs /root/reference). Test infrastructure, never shipped.

Imports the UNMODIFIED reference modules (with the 2-symbol omegaconf stub), checks oracle/unet_oracle.py against them
on identical weights and inputs, and writes the reference's own outputs as golden fixtures to tests/golden/*.npz.

    python oracle/make_golden.py            # validate + (re)write fixtures
    python oracle/make_golden.py --full     # additionally validate the full 865.9 M-parameter config at 16x32 latent
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, os.path.join(HERE, "omegaconf_stub"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from oracle import unet_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def ref_unet(cfg, sd, multiview=None):
    if multiview is None:
        from ldm.modules.diffusionmodules.openaimodel import UNetModel
        m = UNetModel(**cfg)
    else:
        from ldm.modules.diffusionmodules.multiview_unet import MultiViewUnetModel
        m = MultiViewUnetModel(**cfg, view_num=multiview[0], concat_target=multiview[1])
    keys = list(m.state_dict().keys())
    assert keys == [n for n, _ in O.unet_spec(cfg)], "oracle parameter walk differs from the reference state dict"
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    m.load_state_dict(sd, strict=True)
    return m.eval()


def inputs(cfg, n, h, w, seed, L=77):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(n, cfg["in_channels"], h, w, generator=g)
    ctx = torch.randn(n, L, cfg["context_dim"], generator=g)
    return x, ctx


def check(name, a, b, tol=2e-5):
    d = (a - b).abs().max().item()
    s = b.abs().max().item()
    print(f"  {name}: max|oracle-ref|={d:.3e} (ref max {s:.3e})")
    assert d <= tol * max(1.0, s), name


def golden_unet():
    cfg = O.SMALL_CFG
    sd = O.make_state_dict(cfg, seed=0)
    m = ref_unet(cfg, sd)
    x, ctx = inputs(cfg, 2, 16, 32, seed=11)
    t = torch.tensor([981, 21], dtype=torch.long)
    taps_ref = {}
    hooks = []
    for i, blk in enumerate(m.input_blocks):
        hooks.append(blk.register_forward_hook(lambda mod, a, o, k=f"input_blocks.{i}": taps_ref.__setitem__(k, o)))
    hooks.append(m.middle_block.register_forward_hook(lambda mod, a, o: taps_ref.__setitem__("middle_block", o)))
    for i, blk in enumerate(m.output_blocks):
        hooks.append(blk.register_forward_hook(lambda mod, a, o, k=f"output_blocks.{i}": taps_ref.__setitem__(k, o)))
    with torch.no_grad():
        y_ref = m(x, t, context=ctx)
        taps = {}
        y_or = O.unet_forward(sd, cfg, x, t, ctx, taps=taps)
    print("UNetModel (small config, 2x9x16x32):")
    check("output", y_or, y_ref)
    for k in taps_ref:
        check(k, taps[k], taps_ref[k])
    keep = ["input_blocks.0", "input_blocks.2", "input_blocks.3", "input_blocks.11", "middle_block", "output_blocks.2",
            "output_blocks.5", "output_blocks.11"]
    np.savez_compressed(os.path.join(GOLD, "unet_small.npz"), x=x.numpy(), t=t.numpy(), context=ctx.numpy(),
                        out=y_ref.numpy(), **{"tap." + k: taps_ref[k].numpy().astype(np.float32) for k in keep})
    # ragged / odd spatial size (H, W multiples of 8 only): exercises partially filled tiles
    x2, ctx2 = inputs(cfg, 1, 24, 40, seed=12, L=50)
    t2 = torch.tensor([501], dtype=torch.long)
    with torch.no_grad():
        y2 = m(x2, t2, context=ctx2)
        check("output 1x9x24x40, L=50", O.unet_forward(sd, cfg, x2, t2, ctx2), y2)
    np.savez_compressed(os.path.join(GOLD, "unet_small_ragged.npz"), x=x2.numpy(), t=t2.numpy(), context=ctx2.numpy(),
                        out=y2.numpy())
    return cfg, sd, m


def golden_multiview():
    cfg = O.SMALL_CFG
    sd = O.make_state_dict(cfg, seed=1)
    for tag, (v, ct), shape in [("v2", (2, False), (4, 16, 16)), ("v3ct", (3, True), (4, 16, 32))]:
        m = ref_unet(cfg, sd, multiview=(v, ct))
        x, ctx = inputs(cfg, shape[0], shape[1], shape[2], seed=21)
        t = torch.full((shape[0],), 741, dtype=torch.long)
        with torch.no_grad():
            y_ref = m(x, t, context=ctx)
            y_or = O.unet_forward(sd, cfg, x, t, ctx, view_num=v, concat_target=ct)
        print(f"MultiViewUnetModel view_num={v} concat_target={ct} {tuple(x.shape)}:")
        check("output", y_or, y_ref)
        np.savez_compressed(os.path.join(GOLD, f"multiview_{tag}.npz"), x=x.numpy(), t=t.numpy(), context=ctx.numpy(),
                            out=y_ref.numpy(), view_num=v, concat_target=int(ct))


class _FakeLDM:
    """What DDIMSampler needs from LatentDiffusion (ddim.py:11-52,342): schedule buffers + apply_model, restating
    ddpm.py:149-170 (register_schedule, linear 0.00085..0.012) and :865-880 / :1348-1351 (hybrid conditioning)."""

    def __init__(self, unet):
        from ldm.modules.diffusionmodules.util import make_beta_schedule
        betas = make_beta_schedule("linear", 1000, linear_start=0.00085, linear_end=0.0120)
        ac = np.cumprod(1.0 - betas, axis=0)
        self.num_timesteps = 1000
        self.betas = torch.tensor(betas, dtype=torch.float32)
        self.alphas_cumprod = torch.tensor(ac, dtype=torch.float32)
        self.alphas_cumprod_prev = torch.tensor(np.append(1.0, ac[:-1]), dtype=torch.float32)
        self.parameterization = "eps"
        self.device = torch.device("cpu")
        self.unet = unet

    def apply_model(self, x_noisy, t, cond):
        xc = torch.cat([x_noisy] + cond["c_concat"], dim=1)
        cc = torch.cat(cond["c_crossattn"], 1)
        return self.unet(xc, t, context=cc)


def golden_ddim(cfg, sd, m):
    import ldm.models.diffusion.ddim as ddim_mod
    from ldm.models.diffusion.ddim import DDIMSampler
    # the reference force-moves schedule buffers to "cuda" (ddim.py:17-21); keep them on the CPU for this run
    DDIMSampler.register_buffer = lambda self, name, attr: setattr(self, name, attr)
    model = _FakeLDM(m)
    g = torch.Generator(device="cpu").manual_seed(31)
    B, H, W = 1, 16, 32
    x_T = torch.randn(B, 4, H, W, generator=g)
    c_concat = torch.randn(B, 5, H, W, generator=g)
    c_concat[:, 0] = (c_concat[:, 0] > 0).float()
    ctx = torch.randn(B, 77, cfg["context_dim"], generator=g)
    uc = torch.randn(1, 77, cfg["context_dim"], generator=g).repeat(B, 1, 1)
    out = {"x_T": x_T.numpy(), "c_concat": c_concat.numpy(), "context": ctx.numpy(), "uc_context": uc.numpy()}
    for tag, eta in [("eta0", 0.0), ("eta1", 1.0)]:
        S = 4
        noises = []
        orig = ddim_mod.noise_like

        def recording_noise_like(shape, device, repeat=False):
            n = orig(shape, device, repeat)
            noises.append(n.clone())
            return n

        ddim_mod.noise_like = recording_noise_like
        torch.manual_seed(1234)
        sampler = DDIMSampler(model)
        cond = {"c_concat": [c_concat], "c_crossattn": [ctx]}
        ucond = {"c_concat": [c_concat], "c_crossattn": [uc]}
        samples, inter = sampler.sample(S, B, (4, H, W), cond, eta=eta, x_T=x_T, verbose=False,
                                        unconditional_guidance_scale=2.5, unconditional_conditioning=ucond,
                                        log_every_t=1)
        ddim_mod.noise_like = orig
        with torch.no_grad():
            s_or = O.ddim_sample(sd, cfg, x_T, c_concat, ctx, uc, S, eta, 2.5, noises)
        print(f"DDIMSampler.sample S={S} eta={eta} cfg=2.5:")
        check("samples", s_or, samples, tol=1e-4)
        out[f"{tag}.samples"] = samples.numpy()
        out[f"{tag}.noises"] = torch.stack(noises).numpy()
        out[f"{tag}.pred_x0_last"] = inter["pred_x0"][-1].numpy()
        out[f"{tag}.n_inter"] = len(inter["x_inter"])
    # the 50-step schedule the README default uses (eta 1.0)
    sampler = DDIMSampler(model)
    sampler.make_schedule(50, ddim_eta=1.0, verbose=False)
    steps, a, ap, sg = O.make_schedule(50, 1.0, O.make_alphas_cumprod())
    ref_a = np.asarray(sampler.ddim_alphas, dtype=np.float64)
    ref_ap = np.asarray(sampler.ddim_alphas_prev, dtype=np.float64)
    ref_sg = np.asarray(sampler.ddim_sigmas, dtype=np.float64)
    assert (steps == sampler.ddim_timesteps).all()
    assert np.abs(a - ref_a).max() < 1e-6 and np.abs(ap - ref_ap).max() < 1e-6 and np.abs(sg - ref_sg).max() < 1e-6
    print("  50-step schedule matches (timesteps, alphas, alphas_prev, sigmas)")
    out["sched50.timesteps"] = np.asarray(sampler.ddim_timesteps)
    out["sched50.alphas"] = ref_a
    out["sched50.alphas_prev"] = ref_ap
    out["sched50.sigmas"] = ref_sg
    out["sched50.sqrt_one_minus_alphas"] = np.asarray(sampler.ddim_sqrt_one_minus_alphas, dtype=np.float64)
    np.savez_compressed(os.path.join(GOLD, "ddim_small.npz"), **out)


def golden_ddim_multi(cfg, sd, m):
    """DDIMSampler.sample with LIST conditioning -> ddim_multi_sampling (ddim.py:104,146-222): two stitched reference
    views, 3 steps, eta 1, cfg 2.5; the per-step noises and Python's `random` state are pinned."""
    import random
    import ldm.models.diffusion.ddim as ddim_mod
    from ldm.models.diffusion.ddim import DDIMSampler
    DDIMSampler.register_buffer = lambda self, name, attr: setattr(self, name, attr)
    model = _FakeLDM(m)
    g = torch.Generator(device="cpu").manual_seed(47)
    B, H, W, V, S = 1, 16, 32, 2, 4
    x_T = [torch.randn(B, 4, H, W, generator=g) for _ in range(V)]
    c_concat = [torch.randn(B, 5, H, W, generator=g) for _ in range(V)]
    for c in c_concat:
        c[:, 0] = (c[:, 0] > 0).float()
    ctx = [torch.randn(B, 77, cfg["context_dim"], generator=g) for _ in range(V)]
    uc = torch.randn(1, 77, cfg["context_dim"], generator=g).repeat(B, 1, 1)
    noises = []
    orig = ddim_mod.noise_like

    def recording_noise_like(shape, device, repeat=False):
        n = orig(shape, device, repeat)
        noises.append(n.clone())
        return n

    ddim_mod.noise_like = recording_noise_like
    torch.manual_seed(4321)
    random.seed(7)
    sampler = DDIMSampler(model)
    cond = [{"c_concat": [c_concat[v]], "c_crossattn": [ctx[v]]} for v in range(V)]
    ucond = [{"c_concat": [c_concat[v]], "c_crossattn": [uc]} for v in range(V)]
    samples, inter = sampler.sample(S, B, (4, H, W), cond, eta=1.0, x_T=[t.clone() for t in x_T], verbose=False,
                                    unconditional_guidance_scale=2.5, unconditional_conditioning=ucond)
    ddim_mod.noise_like = orig
    assert inter == {} and len(noises) == S * V
    print(f"DDIMSampler.sample (list conditioning -> ddim_multi_sampling) V={V} S={S}: samples {tuple(samples.shape)}")
    out = {"samples": samples.numpy(), "noises": torch.stack(noises).numpy(), "uc_context": uc.numpy(), "S": S, "V": V,
           "random_seed": 7}
    for v in range(V):
        out[f"x_T{v}"] = x_T[v].numpy()
        out[f"c_concat{v}"] = c_concat[v].numpy()
        out[f"context{v}"] = ctx[v].numpy()
    np.savez_compressed(os.path.join(GOLD, "ddim_multi_small.npz"), **out)


def validate_full():
    cfg = O.DEFAULT_CFG
    t0 = time.time()
    sd = O.make_state_dict(cfg, seed=0)
    m = ref_unet(cfg, sd)
    x, ctx = inputs(cfg, 1, 16, 32, seed=41)
    t = torch.tensor([981], dtype=torch.long)
    with torch.no_grad():
        y_ref = m(x, t, context=ctx)
        y_or = O.unet_forward(sd, cfg, x, t, ctx)
    print(f"UNetModel (full ref_inpainting.yaml config, 865.9 M params, 1x9x16x32) [{time.time() - t0:.0f}s]:")
    check("output", y_or, y_ref)
    np.savez_compressed(os.path.join(GOLD, "unet_full_16x32.npz"), x=x.numpy(), t=t.numpy(), context=ctx.numpy(),
                        out=y_ref.numpy())


def _ref_nvs_class():
    """The reference NVSUnetModel (inpainting_ldm/NVS_ldm.py:22-104), UNMODIFIED: its module cannot be imported here
    (it pulls pytorch_lightning, torchmetrics, skimage and the datasets at import time), so the class definition alone
    is extracted from the reference file with `ast` and executed against the reference's own UNetModel / Downsample /
    Upsample / timestep_embedding. Nothing of it is copied into this repository."""
    import ast
    import torch.nn as nn
    from ldm.modules.diffusionmodules.openaimodel import Downsample, UNetModel, Upsample, timestep_embedding
    path = os.path.join(REF, "inpainting_ldm", "NVS_ldm.py")
    src = open(path).read()
    tree = ast.parse(src)
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "NVSUnetModel")
    mod = ast.Module(body=[node], type_ignores=[])
    ns = {"UNetModel": UNetModel, "Downsample": Downsample, "Upsample": Upsample, "nn": nn, "torch": torch,
          "timestep_embedding": timestep_embedding}
    exec(compile(mod, path, "exec"), ns)
    return ns["NVSUnetModel"]


def golden_vae():
    """First-stage decode: the unmodified reference Decoder (ldm/modules/diffusionmodules/model.py:547-653) + the
    post_quant_conv of AutoencoderKL.decode (autoencoder.py:87-90; that class itself needs pytorch_lightning, its decode
    is `decoder(post_quant_conv(z))`), small config (ch = 32), latent 2x4x16x32 -> image 2x3x128x256."""
    from oracle import vae_oracle as V
    from ldm.modules.diffusionmodules.model import Decoder
    cfg = V.SMALL_CFG
    sd = V.make_state_dict(cfg, seed=0)
    dd = {k: v for k, v in cfg.items() if k != "embed_dim"}
    dec = Decoder(**dd)
    keys = ["decoder." + k for k in dec.state_dict().keys()]
    spec = V.decoder_spec(cfg)
    assert keys == [n for n, _ in spec if n.startswith("decoder.")], "oracle decoder walk differs from the reference"
    for n, shp in spec:
        if n.startswith("decoder."):
            assert tuple(dec.state_dict()[n[len("decoder."):]].shape) == tuple(shp), n
    dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")}, strict=True)
    dec.eval()
    pq = torch.nn.Conv2d(cfg["embed_dim"], cfg["z_channels"], 1)
    pq.load_state_dict({"weight": sd["post_quant_conv.weight"], "bias": sd["post_quant_conv.bias"]})
    g = torch.Generator(device="cpu").manual_seed(71)
    z = torch.randn(2, 4, 16, 32, generator=g) * V.SCALE_FACTOR * 4.0     # latents as the sampler returns them
    taps_ref = {}
    hooks = [dec.mid.block_2.register_forward_hook(lambda m, a, o: taps_ref.__setitem__("mid", o))]
    with torch.no_grad():
        y_ref = dec(pq(z / V.SCALE_FACTOR))
        taps = {}
        y_or = V.decode(sd, cfg, z, scale_factor=V.SCALE_FACTOR, taps=taps)
    print("Decoder + post_quant_conv (small config, 2x4x16x32 -> 2x3x128x256):")
    check("image", y_or, y_ref)
    check("mid", taps["mid"], taps_ref["mid"])
    for h in hooks:
        h.remove()
    np.savez_compressed(os.path.join(GOLD, "vae_small.npz"), z=z.numpy(), out=y_ref.numpy(),
                        mid=taps_ref["mid"].numpy(), scale_factor=V.SCALE_FACTOR)
    # the full SD2 decoder (ch = 128, 49.5 M parameters) at a small latent: walk + numerics
    cfg = V.DEFAULT_CFG
    sd = V.make_state_dict(cfg, seed=0)
    dec = Decoder(**{k: v for k, v in cfg.items() if k != "embed_dim"})
    assert ["decoder." + k for k in dec.state_dict().keys()] == [n for n, _ in V.decoder_spec(cfg) if n.startswith("decoder.")]
    dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")}, strict=True)
    dec.eval()
    z = torch.randn(1, 4, 8, 16, generator=g) * V.SCALE_FACTOR * 4.0
    with torch.no_grad():
        zq = torch.nn.functional.conv2d(z / V.SCALE_FACTOR, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
        y_ref = dec(zq)
        y_or = V.decode(sd, cfg, z, scale_factor=V.SCALE_FACTOR)
    print("Decoder (full SD2 VAE config, 1x4x8x16 -> 1x3x64x128):")
    check("image", y_or, y_ref)
    np.savez_compressed(os.path.join(GOLD, "vae_full_8x16.npz"), z=z.numpy(), out=y_ref.numpy(),
                        scale_factor=V.SCALE_FACTOR)


def golden_nvs():
    """NVSUnetModel(use_sep=True) with and without c_input, full config (the reference hard-codes the separator channel
    list for model_channels = 320) at a 16x32 latent."""
    cfg = O.DEFAULT_CFG
    t0 = time.time()
    NVS = _ref_nvs_class()
    sd = O.make_state_dict(cfg, seed=3)
    sd.update(O.make_sep_tokens(cfg, seed=3))
    m = NVS(**cfg, use_sep=True)
    assert sorted(k for k in m.state_dict() if k.startswith("sep_token.")) == sorted(
        f"sep_token.{c}" for c in O.sep_channels(cfg)), "oracle separator channel walk differs from the reference list"
    m.load_state_dict(sd, strict=True)
    m.eval()
    x, ctx = inputs(cfg, 2, 16, 32, seed=61)
    t = torch.tensor([981, 301], dtype=torch.long)
    g = torch.Generator(device="cpu").manual_seed(62)
    c_half = torch.randn(2, cfg["model_channels"], 16, 17, generator=g) * 0.5   # right part of the 33-wide canvas
    c_full = torch.randn(2, cfg["model_channels"], 16, 33, generator=g) * 0.5
    out = {"x": x.numpy(), "t": t.numpy(), "context": ctx.numpy(), "c_input_half": c_half.numpy(),
           "c_input_full": c_full.numpy(), "seed": 3}
    print(f"NVSUnetModel use_sep=True (full config, 2x9x16x32):")
    for tag, ci in [("sep", None), ("sep_cin_half", c_half), ("sep_cin_full", c_full)]:
        with torch.no_grad():
            y_ref = m(x, t, context=ctx, c_input=None if ci is None else ci.clone())
            y_or = O.nvs_unet_forward(sd, cfg, x, t, ctx, use_sep=True, c_input=ci)
        check(tag, y_or, y_ref)
        out[f"out.{tag}"] = y_ref.numpy()
    # use_sep=False + c_input (the plain UNet with input refinement)
    m2 = NVS(**cfg, use_sep=False)
    m2.load_state_dict({k: v for k, v in sd.items() if not k.startswith("sep_token.")}, strict=True)
    m2.eval()
    c_half2 = torch.randn(2, cfg["model_channels"], 16, 16, generator=g) * 0.5
    with torch.no_grad():
        y_ref = m2(x, t, context=ctx, c_input=c_half2.clone())
        y_or = O.nvs_unet_forward(sd, cfg, x, t, ctx, use_sep=False, c_input=c_half2)
    check("nosep_cin_half", y_or, y_ref)
    out["c_input_half_nosep"] = c_half2.numpy()
    out["out.nosep_cin_half"] = y_ref.numpy()
    np.savez_compressed(os.path.join(GOLD, "nvs_full_16x32.npz"), **out)
    print(f"  [{time.time() - t0:.0f}s]")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--only-multi", action="store_true", help="regenerate only tests/golden/ddim_multi_small.npz")
    ap.add_argument("--only-nvs", action="store_true", help="regenerate only tests/golden/nvs_full_16x32.npz")
    ap.add_argument("--only-vae", action="store_true", help="regenerate only tests/golden/vae_*.npz")
    args = ap.parse_args()
    assert os.path.isdir(REF), "the reference tree is only available in the build container"
    os.makedirs(GOLD, exist_ok=True)
    torch.set_grad_enabled(False)
    if args.only_nvs:
        golden_nvs()
        sys.exit(0)
    if args.only_vae:
        golden_vae()
        sys.exit(0)
    if args.only_multi:
        cfg = O.SMALL_CFG
        sd = O.make_state_dict(cfg, seed=0)
        golden_ddim_multi(cfg, sd, ref_unet(cfg, sd))
        sys.exit(0)
    cfg, sd, m = golden_unet()
    golden_multiview()
    golden_ddim(cfg, sd, m)
    golden_ddim_multi(cfg, sd, m)
    golden_vae()
    if args.full:
        validate_full()
        golden_nvs()
    print("golden fixtures written to", GOLD)
