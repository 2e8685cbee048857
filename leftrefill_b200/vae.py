"""Drop-in first-stage DECODER for LeftRefill's latent pipeline.

Mirrors the reference module tree so SD checkpoints (`first_stage_model.decoder.*`, `first_stage_model.post_quant_conv.*`)
load unchanged (ldm/models/autoencoder.py:14-90 AutoencoderKL; ldm/modules/diffusionmodules/model.py:41-204 Normalize /
Upsample / ResnetBlock / AttnBlock, :547-653 Decoder). The modules only OWN parameters; `Decoder.forward` /
`AutoencoderKL.decode` run in liblr_b200.so (lr_vae_* in include/lr_b200.h): the same tcgen05 conv / GEMM and GroupNorm
kernels as the UNet, the mid-block attention (single head, d = 512) as two GEMMs around a row-softmax kernel.
There is no PyTorch fallback: without the library or a CUDA device `decode` raises.

Only the decode half of the autoencoder is native (it is what follows the 50 DDIM steps for every batch); `encode` is
not implemented here (the encoder runs once per batch on the inputs and stays with the reference implementation).
"""
import ctypes

import torch
import torch.nn as nn

from . import _native as N


def Normalize(in_channels, num_groups=32):
    return nn.GroupNorm(num_groups=num_groups, num_channels=in_channels, eps=1e-6, affine=True)


class Upsample(nn.Module):
    def __init__(self, in_channels, with_conv):
        super().__init__()
        if not with_conv:
            raise NotImplementedError("only resamp_with_conv=True (the SD VAE) is implemented")
        self.with_conv = with_conv
        self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1)


class ResnetBlock(nn.Module):
    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout=0.0, temb_channels=0):
        super().__init__()
        if conv_shortcut or temb_channels:
            raise NotImplementedError("the VAE decoder uses nin_shortcut and no time embedding")
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = Normalize(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.norm2 = Normalize(out_channels)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        if in_channels != out_channels:
            self.nin_shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=1, padding=0)


class AttnBlock(nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.norm = Normalize(in_channels)
        self.q = nn.Conv2d(in_channels, in_channels, kernel_size=1)
        self.k = nn.Conv2d(in_channels, in_channels, kernel_size=1)
        self.v = nn.Conv2d(in_channels, in_channels, kernel_size=1)
        self.proj_out = nn.Conv2d(in_channels, in_channels, kernel_size=1)


class _Engine:
    def __init__(self, h):
        self.h = h

    def __del__(self):
        try:
            N.lib().lr_vae_destroy(self.h)
        except Exception:  # noqa: BLE001
            pass


class Decoder(nn.Module):
    """model.py:547-653 constructor signature. forward(z) expects the POST-quant-conv latent only through
    AutoencoderKL.decode below; stand-alone use of this class decodes with an identity post_quant_conv."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, give_pre_end=False, tanh_out=False,
                 use_linear_attn=False, attn_type="vanilla", **ignorekwargs):
        super().__init__()
        if list(attn_resolutions) or give_pre_end or tanh_out or use_linear_attn or attn_type not in ("vanilla",
                                                                                                   "vanilla-xformers"):
            raise NotImplementedError("only the SD VAE decoder layout (mid-block attention only, vanilla attention) "
                                      "is implemented")
        self.ch, self.out_ch, self.ch_mult = ch, out_ch, list(ch_mult)
        self.num_resolutions, self.num_res_blocks = len(ch_mult), num_res_blocks
        self.resolution, self.in_channels, self.z_channels = resolution, in_channels, z_channels
        block_in = ch * ch_mult[-1]
        curr_res = resolution // 2 ** (self.num_resolutions - 1)
        self.z_shape = (1, z_channels, curr_res, curr_res)
        self.conv_in = nn.Conv2d(z_channels, block_in, kernel_size=3, stride=1, padding=1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, dropout=dropout)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, dropout=dropout)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block = nn.ModuleList()
            block_out = ch * ch_mult[i_level]
            for _ in range(num_res_blocks + 1):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, dropout=dropout))
                block_in = block_out
            up = nn.Module()
            up.block = block
            up.attn = nn.ModuleList()
            if i_level != 0:
                up.upsample = Upsample(block_in, resamp_with_conv)
            self.up.insert(0, up)
        self.norm_out = Normalize(block_in)
        self.conv_out = nn.Conv2d(block_in, out_ch, kernel_size=3, stride=1, padding=1)


class AutoencoderKL(nn.Module):
    """Decode half of ldm/models/autoencoder.py:14-90: `decode(z) = decoder(post_quant_conv(z))`, native.

    AutoencoderKL(ddconfig=..., embed_dim=4): same constructor keys as the yaml (lossconfig etc. are accepted and
    ignored). `decode(z, z_scale=1.0)` takes the latent [N, embed_dim, H, W]; LatentDiffusion.decode_first_stage's
    `1 / scale_factor` (ddpm.py:842) can be passed as z_scale and is folded into the input kernel."""

    def __init__(self, ddconfig, embed_dim, lossconfig=None, ckpt_path=None, ignore_keys=(), image_key="image",
                 colorize_nlabels=None, monitor=None, ema_decay=None, learn_logvar=False, **unused):
        super().__init__()
        dd = dict(ddconfig)
        self.embed_dim = embed_dim
        self.decoder = Decoder(**dd)
        assert dd.get("double_z", True)
        self.post_quant_conv = nn.Conv2d(embed_dim, dd["z_channels"], 1)
        self._engine = None
        self._synced = {}

    # parameters the native decoder knows (an encoder / quant_conv loaded from a full checkpoint would be extra keys)
    def _native_params(self):
        for name, p in self.named_parameters():
            if name.startswith(("decoder.", "post_quant_conv.")):
                yield name, p

    def _cfg(self):
        d = self.decoder
        cfg = N.VaeCfg()
        cfg.ch, cfg.out_ch, cfg.num_levels = d.ch, d.out_ch, d.num_resolutions
        for i, m in enumerate(d.ch_mult):
            cfg.ch_mult[i] = int(m)
        cfg.num_res_blocks, cfg.z_channels, cfg.embed_dim = d.num_res_blocks, d.z_channels, self.embed_dim
        return cfg

    def engine(self):
        if self._engine is None:
            h = ctypes.c_void_p()
            cfg = self._cfg()
            N.check(N.lib().lr_vae_create(ctypes.byref(cfg), ctypes.byref(h)), "lr_vae_create")
            self._engine = _Engine(h)
        return self._engine.h

    def engine_weight_names(self):
        L, h = N.lib(), self.engine()
        return [L.lr_vae_weight_name(h, i).decode() for i in range(L.lr_vae_num_weights(h))]

    def __getstate__(self):
        st = self.__dict__.copy()
        st["_engine"], st["_synced"] = None, {}
        return st

    def _apply(self, fn, *args, **kwargs):
        self._synced = {}
        return super()._apply(fn, *args, **kwargs)

    def invalidate_weights(self):
        self._synced = {}

    def sync_weights(self, force=False):
        L, h = N.lib(), self.engine()
        stream = N.current_stream()
        keep = []
        if force:
            self._synced = {}
        for name, p in self._native_params():
            sig = (p.data_ptr(), p._version, p.dtype)
            if self._synced.get(name) == sig:
                continue
            if not p.is_cuda:
                raise N.LRError("AutoencoderKL parameters must live on a CUDA device (call model.to('cuda'))")
            t = p.detach()
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.float().contiguous()
                keep.append(t)
            shape = (ctypes.c_int64 * t.dim())(*t.shape)
            N.check(L.lr_vae_set_weight(h, name.encode(), N.ptr(t), shape, t.dim(), stream), f"lr_vae_set_weight({name})")
            self._synced[name] = sig
        if keep:
            torch.cuda.current_stream().synchronize()

    @torch.no_grad()
    def decode(self, z, z_scale=1.0):
        if not z.is_cuda:
            raise N.LRError("leftrefill_b200.AutoencoderKL.decode runs on CUDA (sm_100a) only; there is no CPU fallback")
        assert z.dim() == 4 and z.shape[1] == self.embed_dim
        with torch.cuda.device(z.device):
            self.sync_weights()
            zf = z.detach().float().contiguous()
            n, _, hh, ww = zf.shape
            f = 2 ** (self.decoder.num_resolutions - 1)
            out = torch.empty(n, self.decoder.out_ch, hh * f, ww * f, dtype=torch.float32, device=z.device)
            N.check(N.lib().lr_vae_decode(self.engine(), N.ptr(zf), float(z_scale), N.ptr(out), n, hh, ww,
                                          N.current_stream()), "lr_vae_decode")
        if torch.is_autocast_enabled():
            return out.half()
        return out

    def last_flops(self):
        return N.lib().lr_vae_last_flops(self.engine())

    def encode(self, x):
        raise NotImplementedError("only the decode half of the first stage is native (SURVEY §8f N2); use the reference "
                                  "AutoencoderKL.encode for the (once per batch) input encoding")
