"""ORACLE — test infrastructure only. Nothing under leftrefill_b200/ may import this file.

A plain-PyTorch fp32 CPU restatement of LeftRefill's DDIM/UNet hot path, written functionally over a flat
state dict so that it shares no code with the product. It follows (reference tree ewrfcas/LeftRefill @ 893c3220):

  unet_spec / unet_forward      ldm/modules/diffusionmodules/openaimodel.py:412-787 (UNetModel.__init__/forward)
  _resblock                     openaimodel.py:254-274 (ResBlock._forward, use_scale_shift_norm=False, no updown)
  _spatial_transformer          ldm/modules/attention.py:393-419
  _transformer_block            attention.py:279-283
  _attention                    attention.py:165-196 (vanilla CrossAttention, fp32 logits)
  _geglu_ff                     attention.py:51-78
  timestep_embedding            ldm/modules/diffusionmodules/util.py:154-174
  make_schedule / ddim_step     ldm/models/diffusion/ddim.py:23-52,304-386; util.py:21-74; ddpm.py:149-170
  multiview self-attention      ldm/modules/multiview_attention.py:431-468 (concat_target False and True)
  nvs_unet_forward              inpainting_ldm/NVS_ldm.py:22-104 (NVSUnetModel: use_sep separator columns, c_input)

Pinning: the reference ships no tests or golden vectors for this path. The oracle is therefore
pinned against the reference ITSELF: oracle/make_golden.py imports the reference modules in the build container,
checks this restatement against them on identical weights/inputs (max |diff| ~1e-6) and commits the reference's
outputs as fixtures under tests/golden/. tests/test_oracle.py re-checks the oracle against those fixtures on CPU.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

DEFAULT_CFG = dict(  # configs/ref_inpainting.yaml:20-36
    image_size=32, in_channels=9, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1],
    num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_head_channels=64, use_spatial_transformer=True,
    use_linear_in_transformer=True, transformer_depth=1, context_dim=1024, legacy=False, use_checkpoint=True)

SMALL_CFG = dict(DEFAULT_CFG, model_channels=64, context_dim=256)  # same topology, 1/25 of the parameters


# ------------------------------------------------------------------------------------------------------------------
# architecture walk: (name, shape) of every parameter, in reference state-dict order
# ------------------------------------------------------------------------------------------------------------------
def _res_spec(p, cin, cout, temb):
    s = [(p + "in_layers.0.weight", (cin,)), (p + "in_layers.0.bias", (cin,)),
         (p + "in_layers.2.weight", (cout, cin, 3, 3)), (p + "in_layers.2.bias", (cout,)),
         (p + "emb_layers.1.weight", (cout, temb)), (p + "emb_layers.1.bias", (cout,)),
         (p + "out_layers.0.weight", (cout,)), (p + "out_layers.0.bias", (cout,)),
         (p + "out_layers.3.weight", (cout, cout, 3, 3)), (p + "out_layers.3.bias", (cout,))]
    if cin != cout:
        s += [(p + "skip_connection.weight", (cout, cin, 1, 1)), (p + "skip_connection.bias", (cout,))]
    return s


def _st_spec(p, C, ctx, depth, use_linear):
    pj = (C, C) if use_linear else (C, C, 1, 1)
    s = [(p + "norm.weight", (C,)), (p + "norm.bias", (C,)), (p + "proj_in.weight", pj), (p + "proj_in.bias", (C,))]
    for d in range(depth):
        b = f"{p}transformer_blocks.{d}."
        s += [(b + "attn1.to_q.weight", (C, C)), (b + "attn1.to_k.weight", (C, C)), (b + "attn1.to_v.weight", (C, C)),
              (b + "attn1.to_out.0.weight", (C, C)), (b + "attn1.to_out.0.bias", (C,)),
              (b + "ff.net.0.proj.weight", (8 * C, C)), (b + "ff.net.0.proj.bias", (8 * C,)),
              (b + "ff.net.2.weight", (C, 4 * C)), (b + "ff.net.2.bias", (C,)),
              (b + "attn2.to_q.weight", (C, C)), (b + "attn2.to_k.weight", (C, ctx)),
              (b + "attn2.to_v.weight", (C, ctx)), (b + "attn2.to_out.0.weight", (C, C)),
              (b + "attn2.to_out.0.bias", (C,)),
              (b + "norm1.weight", (C,)), (b + "norm1.bias", (C,)), (b + "norm2.weight", (C,)),
              (b + "norm2.bias", (C,)), (b + "norm3.weight", (C,)), (b + "norm3.bias", (C,))]
    s += [(p + "proj_out.weight", pj), (p + "proj_out.bias", (C,))]
    return s


def _levels(cfg):
    mult = list(cfg["channel_mult"])
    nrb = cfg["num_res_blocks"]
    nrb = [nrb] * len(mult) if isinstance(nrb, int) else list(nrb)
    return mult, nrb


def unet_layout(cfg):
    """Block structure: list of input blocks / middle / output blocks, each a list of (kind, prefix, meta)."""
    mc = cfg["model_channels"]
    temb = 4 * mc
    mult, nrb = _levels(cfg)
    attn = set(cfg["attention_resolutions"])
    inp = [[("conv_in", "input_blocks.0.0.", (cfg["in_channels"], mc))]]
    chans = [mc]
    ch, ds, ib = mc, 1, 1
    for level, m in enumerate(mult):
        for _ in range(nrb[level]):
            p = f"input_blocks.{ib}."
            blk = [("res", p + "0.", (ch, m * mc, temb))]
            ch = m * mc
            if ds in attn:
                blk.append(("st", p + "1.", (ch,)))
            inp.append(blk)
            chans.append(ch)
            ib += 1
        if level != len(mult) - 1:
            inp.append([("down", f"input_blocks.{ib}.0.op.", (ch, ch))])
            chans.append(ch)
            ds *= 2
            ib += 1
    mid = [("res", "middle_block.0.", (ch, ch, temb)), ("st", "middle_block.1.", (ch,)),
           ("res", "middle_block.2.", (ch, ch, temb))]
    out = []
    ob = 0
    for level in reversed(range(len(mult))):
        m = mult[level]
        for i in range(nrb[level] + 1):
            ich = chans.pop()
            p = f"output_blocks.{ob}."
            blk = [("res", p + "0.", (ch + ich, mc * m, temb))]
            ch = mc * m
            sub = 1
            if ds in attn:
                blk.append(("st", f"{p}{sub}.", (ch,)))
                sub += 1
            if level and i == nrb[level]:
                blk.append(("up", f"{p}{sub}.conv.", (ch, ch)))
                ds //= 2
            out.append(blk)
            ob += 1
    return inp, mid, out, ch


def unet_spec(cfg):
    mc = cfg["model_channels"]
    temb = 4 * mc
    ctx, depth = cfg["context_dim"], cfg.get("transformer_depth", 1)
    use_linear = cfg.get("use_linear_in_transformer", False)
    spec = [("time_embed.0.weight", (temb, mc)), ("time_embed.0.bias", (temb,)),
            ("time_embed.2.weight", (temb, temb)), ("time_embed.2.bias", (temb,))]
    inp, mid, out, ch = unet_layout(cfg)

    def walk(blocks):
        s = []
        for blk in blocks:
            for kind, p, meta in blk:
                if kind == "res":
                    s += _res_spec(p, *meta)
                elif kind == "st":
                    s += _st_spec(p, meta[0], ctx, depth, use_linear)
                else:  # conv_in / down / up
                    s += [(p + "weight", (meta[1], meta[0], 3, 3)), (p + "bias", (meta[1],))]
        return s

    spec += walk(inp) + walk([mid]) + walk(out)
    spec += [("out.0.weight", (ch,)), ("out.0.bias", (ch,)),
             ("out.2.weight", (cfg["out_channels"], mc, 3, 3)), ("out.2.bias", (cfg["out_channels"],))]
    return spec


def sep_channels(cfg):
    """Channel counts that get a separator token (NVS_ldm.py:27 hard-codes [9, 320, 640, 1280, 2560, 1920, 960] for
    model_channels = 320): the input channels of every block that does not end in a Downsample / Upsample."""
    inp, mid, out, _ = unet_layout(cfg)
    order = []
    for blk in inp + [mid] + out:
        if blk[-1][0] in ("down", "up"):
            continue
        c = blk[0][2][0]
        if c not in order:
            order.append(c)
    return order


def make_sep_tokens(cfg, seed=0):
    """Deterministic stand-ins for the learned `sep_token.<channels>` parameters (NVS_ldm.py:28-29: randn init)."""
    sd = {}
    for i, c in enumerate(sep_channels(cfg)):
        g = torch.Generator(device="cpu").manual_seed(seed * 7919 + 100 + i)
        sd[f"sep_token.{c}"] = torch.randn(c, generator=g)
    return sd


def make_state_dict(cfg, seed=0, dtype=torch.float32):
    """Deterministic synthetic weights (no checkpoint is available offline). Variance-preserving scales so that
    activations stay O(1) through ~60 layers; the reference's zero-initialised tensors (zero_module) are drawn like
    every other tensor, otherwise the network output is identically its bias."""
    sd = {}
    for i, (name, shape) in enumerate(unet_spec(cfg)):
        g = torch.Generator(device="cpu").manual_seed(seed * 1000003 + i)
        if len(shape) == 1:
            t = torch.randn(shape, generator=g) * 0.05
            if name.endswith("weight"):  # norm gains
                t = t * 2.0 + 1.0
        else:
            fan_in = int(np.prod(shape[1:]))
            t = torch.randn(shape, generator=g) * (1.0 / math.sqrt(fan_in))
        sd[name] = t.to(dtype)
    return sd


# ------------------------------------------------------------------------------------------------------------------
# forward
# ------------------------------------------------------------------------------------------------------------------
def timestep_embedding(t, dim, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _resblock(sd, p, x, emb):
    h = F.group_norm(x, 32, sd[p + "in_layers.0.weight"], sd[p + "in_layers.0.bias"], 1e-5)
    h = F.conv2d(F.silu(h), sd[p + "in_layers.2.weight"], sd[p + "in_layers.2.bias"], padding=1)
    e = F.linear(F.silu(emb), sd[p + "emb_layers.1.weight"], sd[p + "emb_layers.1.bias"])
    h = h + e[:, :, None, None]
    h = F.group_norm(h, 32, sd[p + "out_layers.0.weight"], sd[p + "out_layers.0.bias"], 1e-5)
    h = F.conv2d(F.silu(h), sd[p + "out_layers.3.weight"], sd[p + "out_layers.3.bias"], padding=1)
    if p + "skip_connection.weight" in sd:
        x = F.conv2d(x, sd[p + "skip_connection.weight"], sd[p + "skip_connection.bias"])
    return x + h


ATTN_IMPL = "vanilla"  # "sdpa": F.scaled_dot_product_attention instead (bench.py's gpu_reference context leg only)


def _attention(sd, p, x, context, heads):
    q = F.linear(x, sd[p + "to_q.weight"])
    k = F.linear(context, sd[p + "to_k.weight"])
    v = F.linear(context, sd[p + "to_v.weight"])
    b, n, c = q.shape
    d = c // heads

    def split(t):
        return t.reshape(b, t.shape[1], heads, d).permute(0, 2, 1, 3)

    q, k, v = split(q), split(k), split(v)
    if ATTN_IMPL == "sdpa":
        o = F.scaled_dot_product_attention(q, k, v)
    else:
        if q.is_cuda and torch.is_autocast_enabled():
            # _ATTN_PRECISION == "fp32" (attention.py:22,175-179): under autocast the logits are computed with autocast
            # disabled on upcast q, k; softmax in fp32; the PV product is back under autocast (fp16)
            with torch.autocast("cuda", enabled=False):
                sim = torch.einsum("bhid,bhjd->bhij", q.float(), k.float()) * (d ** -0.5)
        else:
            sim = torch.einsum("bhid,bhjd->bhij", q, k) * (d ** -0.5)
        o = torch.einsum("bhij,bhjd->bhid", sim.softmax(dim=-1), v)
    o = o.permute(0, 2, 1, 3).reshape(b, n, c)
    return F.linear(o, sd[p + "to_out.0.weight"], sd[p + "to_out.0.bias"])


def _multiview_rearrange(x, view_num, concat_target):
    """multiview_attention.py:436-448: regroup tokens so that all views of a sample attend to each other.
    Returns (sequence [b, T, c], inverse function)."""
    bv, hw, c = x.shape
    if not concat_target:
        v = view_num
        return x.reshape(bv // v, v * hw, c), (lambda s: s.reshape(bv, hw, c))    # '(b v) hw c -> b (v hw) c'
    v = view_num - 1                                            # rows are stitched [ref_i | target] canvases
    side = int(math.sqrt(hw / 2))
    xn = x.reshape(bv // v, v, side, 2 * side, c)
    seq = torch.cat([xn[:, 0:1, :, side:, :], xn[:, :, :, :side, :]], dim=1)    # [target(row 0), ref_1..ref_v]
    seq = seq.reshape(bv // v, (v + 1) * side * side, c)

    def inverse(s):                                             # :456-460, target block broadcast to every row
        s = s.reshape(bv // v, v + 1, side, side, c)
        out = torch.zeros_like(xn)
        out[:, :, :, side:, :] = s[:, 0:1]
        out[:, :, :, :side, :] = s[:, 1:]
        return out.reshape(bv, hw, c)

    return seq, inverse


def _transformer_block(sd, p, x, context, heads, view_num=1, concat_target=False):
    C = x.shape[-1]
    inverse = None
    if view_num > 1:
        x, inverse = _multiview_rearrange(x, view_num, concat_target)
    n1 = F.layer_norm(x, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
    x = _attention(sd, p + "attn1.", n1, n1, heads) + x
    if inverse is not None:
        x = inverse(x)
    n2 = F.layer_norm(x, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
    x = _attention(sd, p + "attn2.", n2, context, heads) + x
    n3 = F.layer_norm(x, (C,), sd[p + "norm3.weight"], sd[p + "norm3.bias"], 1e-5)
    hgate = F.linear(n3, sd[p + "ff.net.0.proj.weight"], sd[p + "ff.net.0.proj.bias"])
    a, g = hgate.chunk(2, dim=-1)
    x = F.linear(a * F.gelu(g), sd[p + "ff.net.2.weight"], sd[p + "ff.net.2.bias"]) + x
    return x


def _spatial_transformer(sd, p, x, context, cfg, view_num=1, concat_target=False):
    b, c, h, w = x.shape
    heads = c // cfg["num_head_channels"]
    use_linear = cfg.get("use_linear_in_transformer", False)
    x_in = x
    x = F.group_norm(x, 32, sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-6)
    if not use_linear:
        x = F.conv2d(x, sd[p + "proj_in.weight"], sd[p + "proj_in.bias"])
    x = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
    if use_linear:
        x = F.linear(x, sd[p + "proj_in.weight"], sd[p + "proj_in.bias"])
    for d in range(cfg.get("transformer_depth", 1)):
        x = _transformer_block(sd, f"{p}transformer_blocks.{d}.", x, context, heads, view_num, concat_target)
    if use_linear:
        x = F.linear(x, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])
    x = x.reshape(b, h, w, c).permute(0, 3, 1, 2)
    if not use_linear:
        x = F.conv2d(x, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])
    return x + x_in


def unet_forward(sd, cfg, x, timesteps, context, view_num=1, concat_target=False, taps=None):
    """UNetModel.forward (openaimodel.py:755-787). `taps`, if a dict, receives the output of every block."""
    inp, mid, out, _ = unet_layout(cfg)
    mc = cfg["model_channels"]
    emb = timestep_embedding(timesteps, mc)
    emb = F.linear(emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])

    def run(blk, h):
        for kind, p, _meta in blk:
            if kind == "res":
                h = _resblock(sd, p, h, emb)
            elif kind == "st":
                h = _spatial_transformer(sd, p, h, context, cfg, view_num, concat_target)
            elif kind == "conv_in":
                h = F.conv2d(h, sd[p + "weight"], sd[p + "bias"], padding=1)
            elif kind == "down":
                h = F.conv2d(h, sd[p + "weight"], sd[p + "bias"], stride=2, padding=1)
            elif kind == "up":
                h = F.interpolate(h, scale_factor=2, mode="nearest")
                h = F.conv2d(h, sd[p + "weight"], sd[p + "bias"], padding=1)
        return h

    hs = []
    h = x.float()
    for i, blk in enumerate(inp):
        h = run(blk, h)
        hs.append(h)
        if taps is not None:
            taps[f"input_blocks.{i}"] = h
    h = run(mid, h)
    if taps is not None:
        taps["middle_block"] = h
    for i, blk in enumerate(out):
        h = torch.cat([h, hs.pop()], dim=1)
        h = run(blk, h)
        if taps is not None:
            taps[f"output_blocks.{i}"] = h
    h = F.group_norm(h, 32, sd["out.0.weight"], sd["out.0.bias"], 1e-5)
    return F.conv2d(F.silu(h), sd["out.2.weight"], sd["out.2.bias"], padding=1)


def nvs_unet_forward(sd, cfg, x, timesteps, context, use_sep=False, c_input=None):
    """NVSUnetModel.forward (inpainting_ldm/NVS_ldm.py:33-104). With use_sep every block that does not end in a
    Downsample / Upsample sees the canvas with one learned column `sep_token.<C>` inserted between its halves (:57-60,
    :74-77, :86-89) and the column is dropped again afterwards (:70-71, :81-82, :93-94); c_input is added to the
    output of input block 0 while the separator is still in place (:64-68)."""
    inp, mid, out, _ = unet_layout(cfg)
    mc = cfg["model_channels"]
    emb = timestep_embedding(timesteps, mc)
    emb = F.linear(emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])

    def run(blk, h):
        for kind, p, _meta in blk:
            if kind == "res":
                h = _resblock(sd, p, h, emb)
            elif kind == "st":
                h = _spatial_transformer(sd, p, h, context, cfg)
            elif kind == "conv_in":
                h = F.conv2d(h, sd[p + "weight"], sd[p + "bias"], padding=1)
            elif kind == "down":
                h = F.conv2d(h, sd[p + "weight"], sd[p + "bias"], stride=2, padding=1)
            elif kind == "up":
                h = F.interpolate(h, scale_factor=2, mode="nearest")
                h = F.conv2d(h, sd[p + "weight"], sd[p + "bias"], padding=1)
        return h

    def sep_blk(blk):
        return use_sep and blk[-1][0] not in ("down", "up")

    def insert(h):
        B, C, H, W = h.shape
        sep = sd[f"sep_token.{C}"].to(h).reshape(1, C, 1, 1).repeat(B, 1, H, 1)
        return torch.cat([h[..., :W // 2], sep, h[..., W // 2:]], dim=-1), W

    def remove(h, W):
        return torch.cat([h[..., :W // 2], h[..., -(W // 2):]], dim=-1)

    hs = []
    h = x.float()
    for i, blk in enumerate(inp):
        W = None
        if sep_blk(blk):
            h, W = insert(h)
        h = run(blk, h)
        if i == 0 and c_input is not None:
            if c_input.shape == h.shape:
                h = h + c_input
            else:
                h = h.clone()
                h[:, :, :, h.shape[-1] // 2:] += c_input
        if W is not None:
            h = remove(h, W)
        hs.append(h)
    W = None
    if use_sep:
        h, W = insert(h)
    h = run(mid, h)
    if W is not None:
        h = remove(h, W)
    for blk in out:
        h = torch.cat([h, hs.pop()], dim=1)
        W = None
        if sep_blk(blk):
            h, W = insert(h)
        h = run(blk, h)
        if W is not None:
            h = remove(h, W)
    h = F.group_norm(h, 32, sd["out.0.weight"], sd["out.0.bias"], 1e-5)
    return F.conv2d(F.silu(h), sd["out.2.weight"], sd["out.2.bias"], padding=1)


# ------------------------------------------------------------------------------------------------------------------
# DDIM schedule and update
# ------------------------------------------------------------------------------------------------------------------
def make_alphas_cumprod(n_timestep=1000, linear_start=0.00085, linear_end=0.0120):
    """ddpm.py:149-170 with make_beta_schedule('linear') (util.py:21-25): float64 math, as the reference."""
    betas = np.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=np.float64) ** 2
    return np.cumprod(1.0 - betas, axis=0)


def make_schedule(S, eta, alphas_cumprod):
    """make_ddim_timesteps('uniform') + make_ddim_sampling_parameters (util.py:46-74)."""
    n = alphas_cumprod.shape[0]
    c = n // S
    steps = np.asarray(list(range(0, n, c))) + 1
    # the reference indexes a float32 torch tensor (ddpm.py register_buffer -> float32) moved to cpu
    ac = alphas_cumprod.astype(np.float32)
    alphas = ac[steps]
    alphas_prev = np.asarray([ac[0]] + ac[steps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return steps, alphas, alphas_prev, sigmas


def ddim_step(x, e_uncond, e_cond, noise, cfg_scale, a_t, a_prev, sigma_t):
    """p_sample_ddim tail (ddim.py:343,359-381), eps-parameterisation."""
    e = e_uncond if e_cond is None else e_uncond + cfg_scale * (e_cond - e_uncond)
    a_t_ = torch.tensor(a_t, dtype=torch.float32)
    a_prev_ = torch.tensor(a_prev, dtype=torch.float32)
    s_ = torch.tensor(sigma_t, dtype=torch.float32)
    sqrt_one_minus_at = torch.tensor(np.sqrt(1.0 - a_t), dtype=torch.float32)
    pred_x0 = (x - sqrt_one_minus_at * e) / a_t_.sqrt()
    dir_xt = (1.0 - a_prev_ - s_ ** 2).sqrt() * e
    x_prev = a_prev_.sqrt() * pred_x0 + dir_xt + s_ * noise
    return x_prev, pred_x0


def ddim_sample(sd, cfg, x_T, c_concat, context, uc_context, S, eta, cfg_scale, noises, view_num=1,
                concat_target=False, trajectory=None, autocast_unet=False):
    """DDIMSampler.sample + DiffusionWrapper 'hybrid' (ddim.py:224-302; ddpm.py:1348-1351) with explicit noises.
    `trajectory`, if a list, receives x after every step. `autocast_unet` runs the UNet call under
    torch.autocast("cuda") (the reference's precision recipe on a GPU, inpainting_ldm/ref_inpainting_ldm.py) with the
    DDIM state in fp32, as the reference does."""
    steps, alphas, alphas_prev, sigmas = make_schedule(S, eta, make_alphas_cumprod())
    x = x_T
    b = x.shape[0]
    dev = x.device
    for i, step in enumerate(np.flip(steps)):
        index = len(steps) - i - 1
        t = torch.full((2 * b,), int(step), dtype=torch.long, device=dev)
        xc = torch.cat([torch.cat([x] * 2), torch.cat([c_concat] * 2)], dim=1)
        cc = torch.cat([uc_context, context])
        if autocast_unet:
            with torch.autocast("cuda"):
                e = unet_forward(sd, cfg, xc, t, cc, view_num, concat_target).float()
        else:
            e = unet_forward(sd, cfg, xc, t, cc, view_num, concat_target)
        e_u, e_c = e.chunk(2)
        x, _ = ddim_step(x, e_u, e_c, noises[i], cfg_scale, float(alphas[index]), float(alphas_prev[index]),
                         float(sigmas[index]))
        if trajectory is not None:
            trajectory.append(x)
    return x


def ddim_multi_sample(sd, cfg, x_T, c_concat, context, uc_context, S, eta, cfg_scale, noises, rng):
    """DDIMSampler.ddim_multi_sampling (ddim.py:146-222): `x_T`, `c_concat`, `context` are lists with one entry per
    reference view, `noises` is the flat list of per-(step, view) noises in call order and `rng` a `random.Random` in the
    state the reference's module-level `random` had (it picks the view whose target half is copied into all canvases
    with `random.shuffle`, ddim.py:205-207). Returns the first canvas, like the reference."""
    steps, alphas, alphas_prev, sigmas = make_schedule(S, eta, make_alphas_cumprod())
    img = [x.clone() for x in x_T]
    b = img[0].shape[0]
    k = 0
    for i, step in enumerate(np.flip(steps)):
        index = len(steps) - i - 1
        t = torch.full((2 * b,), int(step), dtype=torch.long)
        rights, new_img = [], []
        for v in range(len(img)):
            xc = torch.cat([torch.cat([img[v]] * 2), torch.cat([c_concat[v]] * 2)], dim=1)
            e = unet_forward(sd, cfg, xc, t, torch.cat([uc_context, context[v]]))
            e_u, e_c = e.chunk(2)
            x, _ = ddim_step(img[v], e_u, e_c, noises[k], cfg_scale, float(alphas[index]), float(alphas_prev[index]),
                             float(sigmas[index]))
            k += 1
            rights.append(x[:, :, :, x.shape[-1] // 2:])
            new_img.append(x)
        rng.shuffle(rights)
        right = rights[0].clone()
        for x in new_img:
            x[:, :, :, right.shape[-1]:] = right
        img = new_img
    return img[0]
