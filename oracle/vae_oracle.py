"""ORACLE — test infrastructure only. Nothing under leftrefill_b200/ may import this file.

Plain-PyTorch fp32 restatement of the first-stage DECODE path of LeftRefill, written functionally over
a flat state dict. It follows (reference tree ewrfcas/LeftRefill @ 893c3220):

  decode                 ldm/models/autoencoder.py:87-90 (post_quant_conv -> decoder) and LatentDiffusion.
                         decode_first_stage (ldm/models/diffusion/ddpm.py:835-843: z / scale_factor first)
  decoder_forward        ldm/modules/diffusionmodules/model.py:547-653 (Decoder.__init__ / forward)
  _resnet_block          model.py:82-150 (ResnetBlock, temb_channels = 0, nin_shortcut when the width changes)
  _attn_block            model.py:153-204 (AttnBlock: GroupNorm, 1x1 q/k/v, single-head softmax(q k^T / sqrt(c)) v, proj_out)
  Normalize / swish      model.py:41-48 (GroupNorm(32, eps 1e-6); x * sigmoid(x))
  Upsample               model.py:51-66 (nearest x2 + 3x3 conv)

Pinning: oracle/make_golden.py --only-vae imports the UNMODIFIED reference Decoder in the build container, checks
this restatement against it on identical weights / latents and commits the reference's output as
tests/golden/vae_small.npz (tests/test_oracle.py re-checks it on CPU).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

DEFAULT_CFG = dict(  # configs/ref_inpainting.yaml:38-58 (first_stage_config.params.ddconfig + embed_dim)
    ch=128, out_ch=3, ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0, in_channels=3,
    resolution=256, z_channels=4, double_z=True, embed_dim=4)
SMALL_CFG = dict(DEFAULT_CFG, ch=32)  # same topology, 1/16 of the parameters
SCALE_FACTOR = 0.18215                # configs/ref_inpainting.yaml scale_factor (SD2)


def _res_spec(p, cin, cout):
    s = [(p + "norm1.weight", (cin,)), (p + "norm1.bias", (cin,)), (p + "conv1.weight", (cout, cin, 3, 3)),
         (p + "conv1.bias", (cout,)), (p + "norm2.weight", (cout,)), (p + "norm2.bias", (cout,)),
         (p + "conv2.weight", (cout, cout, 3, 3)), (p + "conv2.bias", (cout,))]
    if cin != cout:
        s += [(p + "nin_shortcut.weight", (cout, cin, 1, 1)), (p + "nin_shortcut.bias", (cout,))]
    return s


def decoder_layout(cfg):
    """[(level, [(cin, cout) per block], has_upsample)] from the lowest resolution up (model.py:586-606)."""
    ch, mult, nrb = cfg["ch"], list(cfg["ch_mult"]), cfg["num_res_blocks"]
    block_in = ch * mult[-1]
    levels = []
    for lvl in reversed(range(len(mult))):
        block_out = ch * mult[lvl]
        blocks = []
        for _ in range(nrb + 1):
            blocks.append((block_in, block_out))
            block_in = block_out
        levels.append((lvl, blocks, lvl != 0))
    return ch * mult[-1], levels, block_in


def decoder_spec(cfg):
    """(name, shape) of every decoder parameter in reference state-dict order, keys relative to AutoencoderKL
    ("decoder.*", then "post_quant_conv.*", autoencoder.py:33-35)."""
    assert not cfg["attn_resolutions"], "only the mid-block attention of the SD VAE is restated"
    top, levels, last = decoder_layout(cfg)
    zc = cfg["z_channels"]
    s = [("decoder.conv_in.weight", (top, zc, 3, 3)), ("decoder.conv_in.bias", (top,))]
    s += _res_spec("decoder.mid.block_1.", top, top)
    a = "decoder.mid.attn_1."
    s += [(a + "norm.weight", (top,)), (a + "norm.bias", (top,))]
    for n in ("q", "k", "v", "proj_out"):
        s += [(a + n + ".weight", (top, top, 1, 1)), (a + n + ".bias", (top,))]
    s += _res_spec("decoder.mid.block_2.", top, top)
    for lvl, blocks, up in sorted(levels):          # ModuleList order is up.0 .. up.N-1 (model.py:606 prepends)
        for i, (cin, cout) in enumerate(blocks):
            s += _res_spec(f"decoder.up.{lvl}.block.{i}.", cin, cout)
        if up:
            c = blocks[-1][1]
            s += [(f"decoder.up.{lvl}.upsample.conv.weight", (c, c, 3, 3)), (f"decoder.up.{lvl}.upsample.conv.bias", (c,))]
    s += [("decoder.norm_out.weight", (last,)), ("decoder.norm_out.bias", (last,)),
          ("decoder.conv_out.weight", (cfg["out_ch"], last, 3, 3)), ("decoder.conv_out.bias", (cfg["out_ch"],))]
    e = cfg["embed_dim"]
    s += [("post_quant_conv.weight", (zc, e, 1, 1)), ("post_quant_conv.bias", (zc,))]
    return s


def make_state_dict(cfg, seed=0):
    """Deterministic synthetic weights (no checkpoint offline), variance preserving."""
    sd = {}
    for i, (name, shape) in enumerate(decoder_spec(cfg)):
        g = torch.Generator(device="cpu").manual_seed(seed * 1000003 + 7000 + i)
        if len(shape) == 1:
            t = torch.randn(shape, generator=g) * 0.05
            if "norm" in name and name.endswith("weight"):
                t = t * 2.0 + 1.0
        else:
            t = torch.randn(shape, generator=g) * (1.0 / math.sqrt(int(np.prod(shape[1:]))))
        sd[name] = t
    return sd


def _swish(x):
    return x * torch.sigmoid(x)


def _norm(sd, p, x):
    return F.group_norm(x, 32, sd[p + "weight"], sd[p + "bias"], 1e-6)


def _resnet_block(sd, p, x):
    h = F.conv2d(_swish(_norm(sd, p + "norm1.", x)), sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1)
    h = F.conv2d(_swish(_norm(sd, p + "norm2.", h)), sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1)
    if p + "nin_shortcut.weight" in sd:
        x = F.conv2d(x, sd[p + "nin_shortcut.weight"], sd[p + "nin_shortcut.bias"])
    return x + h


def _attn_block(sd, p, x):
    h_ = _norm(sd, p + "norm.", x)
    q = F.conv2d(h_, sd[p + "q.weight"], sd[p + "q.bias"])
    k = F.conv2d(h_, sd[p + "k.weight"], sd[p + "k.bias"])
    v = F.conv2d(h_, sd[p + "v.weight"], sd[p + "v.bias"])
    b, c, h, w = q.shape
    q = q.reshape(b, c, h * w).permute(0, 2, 1)
    k = k.reshape(b, c, h * w)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, h * w)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)
    return x + F.conv2d(h_, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])


def decoder_forward(sd, cfg, z, taps=None):
    """Decoder.forward (model.py:608-653) on the post-quant latent."""
    _, levels, _ = decoder_layout(cfg)
    p = "decoder."
    h = F.conv2d(z, sd[p + "conv_in.weight"], sd[p + "conv_in.bias"], padding=1)
    h = _resnet_block(sd, p + "mid.block_1.", h)
    h = _attn_block(sd, p + "mid.attn_1.", h)
    h = _resnet_block(sd, p + "mid.block_2.", h)
    if taps is not None:
        taps["mid"] = h
    for lvl, blocks, up in levels:
        for i in range(len(blocks)):
            h = _resnet_block(sd, f"{p}up.{lvl}.block.{i}.", h)
        if up:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, sd[f"{p}up.{lvl}.upsample.conv.weight"], sd[f"{p}up.{lvl}.upsample.conv.bias"], padding=1)
        if taps is not None:
            taps[f"up.{lvl}"] = h
    h = _swish(_norm(sd, p + "norm_out.", h))
    return F.conv2d(h, sd[p + "conv_out.weight"], sd[p + "conv_out.bias"], padding=1)


def decode(sd, cfg, z, scale_factor=None, taps=None):
    """decode_first_stage (ddpm.py:842-843: z / scale_factor) + AutoencoderKL.decode (autoencoder.py:87-90)."""
    if scale_factor is not None:
        z = (1.0 / scale_factor) * z
    z = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    return decoder_forward(sd, cfg, z, taps)
