"""Drop-in `UNetModel` (and its building blocks) for LeftRefill's SD2-inpainting UNet.

Call surface and parameter names follow the reference ldm/modules/diffusionmodules/openaimodel.py (UNetModel
:412-787, ResBlock :162-274, Upsample :90-118, Downsample :133-159, TimestepEmbedSequential :73-87) so that
`instantiate_from_config`, SD2 checkpoints (`model.diffusion_model.*`, 686 tensors) and `torch_init_model`
(test_inpainting.py:26-53) work unchanged. The modules here only OWN parameters; all arithmetic of
`UNetModel.forward` runs in liblr_b200.so (hand-written sm_100a CUDA) through the C ABI in include/lr_b200.h.
There is no PyTorch fallback: without the library or without a CUDA device `forward` raises.
"""
import ctypes

import torch
import torch.nn as nn

from . import _native as N
from . import ops
from .attention import SpatialTransformer, zero_module


class GroupNorm32(nn.GroupNorm):
    """Parameter holder for `normalization(channels)` (util.py:202-219): 32 groups, eps 1e-5, fp32 statistics."""

    def forward(self, x):
        return _run_nchw(lambda t: ops.groupnorm(t, self.weight.float(), self.bias.float(), self.eps, silu=False,
                                                 groups=self.num_groups), x)


def normalization(channels):
    return GroupNorm32(32, channels)


def _run_nchw(fn, x):
    """Stand-alone use of a block: NCHW tensor in/out around an NHWC fp16 native op."""
    y = ops.to_nchw_f32(fn(ops.to_nhwc_f16(x)))
    return y.to(x.dtype) if x.dtype != torch.float32 else y


class TimestepBlock(nn.Module):
    """Marker base class: forward(x, emb)."""


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    def forward(self, x, emb, context=None, **kwargs):
        for layer in self:
            if isinstance(layer, TimestepBlock):
                x = layer(x, emb)
            elif isinstance(layer, SpatialTransformer):
                x = layer(x, context)
            else:
                x = layer(x)
        return x


class Upsample(nn.Module):
    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        assert dims == 2, "only 2-D UNets are supported"
        self.channels, self.out_channels, self.use_conv, self.dims = channels, out_channels or channels, use_conv, dims
        if use_conv:
            self.conv = nn.Conv2d(self.channels, self.out_channels, 3, padding=padding)

    def forward(self, x):
        assert x.shape[1] == self.channels

        def f(t):
            n, h, w, c = t.shape
            t = t[:, :, None, :, None, :].expand(n, h, 2, w, 2, c).reshape(n, 2 * h, 2 * w, c).contiguous()
            if self.use_conv:
                t = ops.conv3x3(t, ops.repack_conv3x3(self.conv.weight), bias=self.conv.bias.float())
            return t

        return _run_nchw(f, x)


class Downsample(nn.Module):
    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        assert dims == 2 and use_conv, "only the learned stride-2 conv downsample is supported"
        self.channels, self.out_channels, self.use_conv, self.dims = channels, out_channels or channels, use_conv, dims
        self.op = nn.Conv2d(self.channels, self.out_channels, 3, stride=2, padding=padding)

    def forward(self, x):
        assert x.shape[1] == self.channels
        return _run_nchw(lambda t: ops.conv3x3(t, ops.repack_conv3x3(self.op.weight), bias=self.op.bias.float(),
                                               stride=2), x)


class ResBlock(TimestepBlock):
    """GroupNorm32+SiLU+conv3x3, + Linear(SiLU(emb)), GroupNorm32+SiLU+conv3x3, + skip (openaimodel.py:254-274)."""

    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_conv=False, use_scale_shift_norm=False,
                 dims=2, use_checkpoint=False, up=False, down=False):
        super().__init__()
        if use_scale_shift_norm or up or down or use_conv or dims != 2:
            raise NotImplementedError("ResBlock: scale-shift norm / resblock_updown / 3x3 skip are not used by any "
                                      "LeftRefill config and are not implemented")
        self.channels, self.emb_channels, self.dropout = channels, emb_channels, dropout
        self.out_channels = out_channels or channels
        self.use_checkpoint = use_checkpoint  # accepted and ignored: inference only
        self.in_layers = nn.Sequential(normalization(channels), nn.SiLU(),
                                       nn.Conv2d(channels, self.out_channels, 3, padding=1))
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, self.out_channels))
        self.out_layers = nn.Sequential(normalization(self.out_channels), nn.SiLU(), nn.Dropout(p=dropout),
                                        zero_module(nn.Conv2d(self.out_channels, self.out_channels, 3, padding=1)))
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        else:
            self.skip_connection = nn.Conv2d(channels, self.out_channels, 1)

    def forward(self, x, emb):
        gn1, conv1 = self.in_layers[0], self.in_layers[2]
        gn2, conv2 = self.out_layers[0], self.out_layers[3]
        lin = self.emb_layers[1]
        e = torch.nn.functional.silu(emb.float()) @ lin.weight.float().t() + lin.bias.float()  # [N, Cout], tiny

        def f(t):
            n, h, w, c = t.shape
            hn = ops.groupnorm(t, gn1.weight.float(), gn1.bias.float(), gn1.eps, silu=True)
            hh = ops.conv3x3(hn, ops.repack_conv3x3(conv1.weight), bias=conv1.bias.float(), bias_img=e.contiguous())
            hn = ops.groupnorm(hh, gn2.weight.float(), gn2.bias.float(), gn2.eps, silu=True)
            if isinstance(self.skip_connection, nn.Identity):
                skip = t
            else:
                sw = ops.repack_linear(self.skip_connection.weight.reshape(self.out_channels, c))
                skip = ops.linear(t.reshape(-1, c), sw, bias=self.skip_connection.bias.float())
                skip = skip.reshape(n, h, w, self.out_channels)
            return ops.conv3x3(hn, ops.repack_conv3x3(conv2.weight), bias=conv2.bias.float(), residual=skip)

        return _run_nchw(f, x)


class UNetModel(nn.Module):
    """SD2-inpainting UNet with the reference constructor signature (openaimodel.py:442-472).

    forward(x [N, in_channels, H, W], timesteps [N], context [N, L, context_dim]) -> [N, out_channels, H, W];
    fp32 in/out at the boundary, fp16 tensor-core arithmetic with fp32 accumulation / norm statistics / softmax inside.
    Safe under torch.no_grad() + torch.autocast("cuda") (autocast does not see the native ops).
    """

    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None,
                 use_checkpoint=False, use_fp16=False, num_heads=-1, num_head_channels=-1, num_heads_upsample=-1,
                 use_scale_shift_norm=False, resblock_updown=False, use_new_attention_order=False,
                 use_spatial_transformer=False, transformer_depth=1, context_dim=None, n_embed=None, legacy=True,
                 disable_self_attentions=None, num_attention_blocks=None, disable_middle_self_attn=False,
                 use_linear_in_transformer=False, view_num=1, concat_target=False, use_sep=False, **unused):
        super().__init__()
        # Options this implementation does not know must not be dropped silently: a truthy unknown kwarg changes what the
        # reference computes (e.g. MultiViewUnetModel's no_rearrange_selfattn, multiview_attention.py:437).
        noop = {"adm_in_channels", "num_attention_blocks", "use_bf16"}
        bad_kw = [k for k, v in unused.items() if k not in noop and v not in (None, False, 0)]
        if bad_kw:
            raise NotImplementedError(f"UNetModel options not implemented (they would change the result): {bad_kw}")
        if not use_spatial_transformer or context_dim is None:
            raise NotImplementedError("only the SpatialTransformer UNet (use_spatial_transformer=True with a "
                                      "context_dim) used by every LeftRefill config is implemented")
        if isinstance(context_dim, (list, tuple)) or type(context_dim).__name__ == "ListConfig":
            context_dim = list(context_dim)
            assert len(set(context_dim)) == 1, "per-depth context dims are not supported"
            context_dim = context_dim[0]
        unsupported = dict(num_classes=num_classes is not None, use_scale_shift_norm=use_scale_shift_norm,
                           resblock_updown=resblock_updown, n_embed=n_embed is not None, dims=dims != 2,
                           disable_self_attentions=disable_self_attentions is not None,
                           num_attention_blocks=num_attention_blocks is not None,
                           disable_middle_self_attn=disable_middle_self_attn, conv_resample=not conv_resample)
        bad = [k for k, v in unsupported.items() if v]
        if bad:
            raise NotImplementedError(f"UNetModel options not used by LeftRefill and not implemented: {bad}")
        if num_head_channels == -1:
            assert num_heads != -1, "Either num_heads or num_head_channels has to be set"
        self.image_size, self.in_channels, self.model_channels = image_size, in_channels, model_channels
        self.out_channels, self.transformer_depth = out_channels, transformer_depth
        self.num_res_blocks = (len(channel_mult) * [num_res_blocks] if isinstance(num_res_blocks, int)
                               else list(num_res_blocks))
        if len(self.num_res_blocks) != len(channel_mult):
            raise ValueError("provide num_res_blocks either as an int (globally constant) or as a list/tuple "
                             "(per-level) with the same length as channel_mult")
        self.attention_resolutions = list(attention_resolutions)
        self.dropout, self.channel_mult, self.conv_resample = dropout, list(channel_mult), conv_resample
        self.num_classes, self.use_checkpoint = num_classes, use_checkpoint
        self.dtype = torch.float16 if use_fp16 else torch.float32
        self.num_heads, self.num_head_channels, self.num_heads_upsample = num_heads, num_head_channels, num_heads_upsample
        self.predict_codebook_ids = False
        self.context_dim, self.use_linear_in_transformer = context_dim, use_linear_in_transformer
        self.view_num, self.concat_target = int(view_num), bool(concat_target)
        self.use_sep = bool(use_sep)
        if self.use_sep and self.view_num > 1:
            raise NotImplementedError("use_sep is an NVSUnetModel option; it is not combined with the multiview UNet")

        mc, temb = model_channels, model_channels * 4

        def heads_of(ch):
            if num_head_channels == -1:
                d = ch // num_heads
            else:
                d = num_head_channels
            if d != 64:
                raise NotImplementedError("the fused attention kernel supports d_head == 64 only "
                                          f"(got {d}); every LeftRefill config uses num_head_channels=64")
            return ch // d, d

        def st(ch):
            nh, d = heads_of(ch)
            return SpatialTransformer(ch, nh, d, depth=transformer_depth, context_dim=context_dim,
                                      use_linear=use_linear_in_transformer, use_checkpoint=use_checkpoint)

        self.time_embed = nn.Sequential(nn.Linear(mc, temb), nn.SiLU(), nn.Linear(temb, temb))
        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(nn.Conv2d(in_channels, mc, 3, padding=1))])
        chans, ch, ds = [mc], mc, 1
        for level, mult in enumerate(channel_mult):
            for _ in range(self.num_res_blocks[level]):
                layers = [ResBlock(ch, temb, dropout, out_channels=mult * mc, use_checkpoint=use_checkpoint)]
                ch = mult * mc
                if ds in self.attention_resolutions:
                    layers.append(st(ch))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(Downsample(ch, True, out_channels=ch)))
                chans.append(ch)
                ds *= 2
        self.middle_block = TimestepEmbedSequential(ResBlock(ch, temb, dropout, use_checkpoint=use_checkpoint), st(ch),
                                                    ResBlock(ch, temb, dropout, use_checkpoint=use_checkpoint))
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(self.num_res_blocks[level] + 1):
                ich = chans.pop()
                layers = [ResBlock(ch + ich, temb, dropout, out_channels=mc * mult, use_checkpoint=use_checkpoint)]
                ch = mc * mult
                if ds in self.attention_resolutions:
                    layers.append(st(ch))
                if level and i == self.num_res_blocks[level]:
                    layers.append(Upsample(ch, True, out_channels=ch))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
        self.out = nn.Sequential(normalization(ch), nn.SiLU(),
                                 zero_module(nn.Conv2d(mc, out_channels, 3, padding=1)))
        if self.use_sep:
            # NVS_ldm.py:24-31: one learned separator token per channel count that enters a non-resampling block. The
            # reference hard-codes [9, 320, 640, 1280, 2560, 1920, 960] (model_channels 320); the same walk for any config:
            self.sep_token = nn.ParameterDict({str(c): nn.Parameter(torch.randn(c)) for c in self._sep_channels()})
        else:
            self.sep_token = None
        self._engine = None
        self._synced = {}

    def _sep_channels(self):
        order = []

        def first_in(block):
            m = block[0]
            if isinstance(m, nn.Conv2d):
                return m.in_channels
            return m.channels

        def sep_block(block):
            return not isinstance(block[-1], (Downsample, Upsample))

        for blk in list(self.input_blocks) + [self.middle_block] + list(self.output_blocks):
            if sep_block(blk) and first_in(blk) not in order:
                order.append(first_in(blk))
        return order

    # engine handles are ctypes pointers: never copied or pickled with the module (the copy re-creates its own lazily)
    def __getstate__(self):
        st = self.__dict__.copy()
        st["_engine"], st["_synced"] = None, {}
        st.pop("_step_graphs", None)
        st.pop("_ctx_keepalive", None)
        return st

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__getstate__().items():
            setattr(new, k, copy.deepcopy(v, memo))
        return new

    # ---- reference API kept for compatibility -------------------------------------------------------------------
    def convert_to_fp16(self):
        """No-op: master weights stay fp32 in PyTorch, the engine always holds its own fp16 copies."""

    def convert_to_fp32(self):
        """No-op (see convert_to_fp16)."""

    # ---- native engine -------------------------------------------------------------------------------------------
    def _cfg(self):
        cfg = N.UNetCfg()
        cfg.in_channels, cfg.model_channels, cfg.out_channels = self.in_channels, self.model_channels, self.out_channels
        cfg.num_levels = len(self.channel_mult)
        for i, m in enumerate(self.channel_mult):
            cfg.channel_mult[i] = int(m)
            cfg.num_res_blocks[i] = int(self.num_res_blocks[i])
        for i, a in enumerate(self.attention_resolutions):
            cfg.attention_ds[i] = int(a)
        cfg.n_attention_ds = len(self.attention_resolutions)
        cfg.num_head_channels = 64
        cfg.transformer_depth = int(self.transformer_depth)
        cfg.context_dim = int(self.context_dim)
        cfg.use_linear_in_transformer = int(bool(self.use_linear_in_transformer))
        cfg.view_num = self.view_num
        cfg.concat_target = int(self.concat_target)
        cfg.use_sep = int(self.use_sep)
        return cfg

    def engine(self):
        if self._engine is None:
            h = ctypes.c_void_p()
            cfg = self._cfg()
            N.check(N.lib().lr_unet_create(ctypes.byref(cfg), ctypes.byref(h)), "lr_unet_create")
            self._engine = _EngineHandle(h)
        return self._engine.h

    def engine_weight_names(self):
        L, h = N.lib(), self.engine()
        return [L.lr_unet_weight_name(h, i).decode() for i in range(L.lr_unet_num_weights(h))]

    def invalidate_weights(self):
        """Forces the next forward / set_context to re-upload every parameter. Needed after writes that PyTorch's version
        counter does not see: `p.data.copy_(...)` / `p.data.mul_(...)` as done by LitEma.copy_to / restore, ema_scope and
        most LoRA-merge utilities (ldm/modules/ema.py)."""
        self._synced = {}

    def _apply(self, fn, *args, **kwargs):  # .to() / .cuda() / .half(): storages move, the next sync re-uploads
        self._synced = {}
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):  # copies through p.data-style paths in some loaders: always re-upload
        self._synced = {}
        return super().load_state_dict(*args, **kwargs)

    def sync_weights(self, force=False):
        """Uploads every parameter whose storage or version changed since the last call (fp32 -> repacked fp16).
        `force=True` (or invalidate_weights()) re-uploads everything."""
        L, h = N.lib(), self.engine()
        stream = N.current_stream()
        keep = []
        if force:
            self._synced = {}
        for name, p in self.named_parameters():
            sig = (p.data_ptr(), p._version, p.dtype)
            if self._synced.get(name) == sig:
                continue
            if not p.is_cuda:
                raise N.LRError("UNetModel parameters must live on a CUDA device (call model.to('cuda'))")
            t = p.detach()
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.float().contiguous()
                keep.append(t)
            shape = (ctypes.c_int64 * t.dim())(*t.shape)
            N.check(L.lr_unet_set_weight(h, name.encode(), N.ptr(t), shape, t.dim(), stream),
                    f"lr_unet_set_weight({name})")
            self._synced[name] = sig
        if keep:
            torch.cuda.current_stream().synchronize()  # temporaries must outlive the repack kernels

    def set_context(self, context):
        """Caches the cross-attention K/V of `context` [N, L, context_dim] for subsequent forward(context=None)."""
        with torch.cuda.device(context.device):
            self.sync_weights()
            c = context.detach().float().contiguous()
            N.check(N.lib().lr_unet_set_context(self.engine(), N.ptr(c), c.shape[0], c.shape[1], N.current_stream()),
                    "lr_unet_set_context")
        self._ctx_keepalive = c

    def set_c_input(self, c_input, unet_width):
        """NVS input refinement (NVS_ldm.py:49,64-68): stages `c_input` [N, model_channels, H, Wc] (or None to clear)
        for the following forwards; `unet_width` is the width of the UNet input x."""
        L, h = N.lib(), self.engine()
        if c_input is None:
            N.check(L.lr_unet_set_c_input(h, None, 0, 0, 0, 0, 0, N.current_stream()), "lr_unet_set_c_input")
            return
        with torch.cuda.device(c_input.device):
            c = c_input.detach().float().contiguous()
            n, ch, hh, wc = c.shape
            N.check(L.lr_unet_set_c_input(h, N.ptr(c), n, ch, hh, wc, int(unet_width), N.current_stream()),
                    "lr_unet_set_c_input")
            self._cin_keepalive = c

    def forward_native(self, x, timesteps, context=None):
        """x fp32 contiguous NCHW CUDA, timesteps int64 [N]; context None -> use the cached K/V."""
        n, c, hh, ww = x.shape
        with torch.cuda.device(x.device):
            out = torch.empty(n, self.out_channels, hh, ww, dtype=torch.float32, device=x.device)
            L = 0
            if context is not None:
                L = context.shape[1]
            N.check(N.lib().lr_unet_forward(self.engine(), N.ptr(x), N.ptr(timesteps), N.ptr(context), L, N.ptr(out), n,
                                            hh, ww, N.current_stream()), "lr_unet_forward")
        return out

    def forward_native_cfg_pair(self, x, timesteps):
        """CFG pair [uncond | cond] sharing x / timesteps: x [B, C, H, W] -> eps [2B, out, H, W]. The 2B contexts (uncond
        first) must have been cached with set_context. Bit-identical to forward_native on the doubled batch."""
        n, c, hh, ww = x.shape
        with torch.cuda.device(x.device):
            out = torch.empty(2 * n, self.out_channels, hh, ww, dtype=torch.float32, device=x.device)
            N.check(N.lib().lr_unet_forward_cfg_pair(self.engine(), N.ptr(x), N.ptr(timesteps), N.ptr(out), n, hh, ww,
                                                     N.current_stream()), "lr_unet_forward_cfg_pair")
        return out

    def forward(self, x, timesteps=None, context=None, y=None, **kwargs):
        assert y is None, "must specify y if and only if the model is class-conditional"
        c_input = kwargs.get("c_input")
        assert timesteps is not None and context is not None
        if not x.is_cuda:
            raise N.LRError("leftrefill_b200.UNetModel runs on CUDA (sm_100a) only; there is no CPU fallback")
        assert x.shape[1] == self.in_channels and x.dim() == 4
        with torch.cuda.device(x.device):
            self.sync_weights()
            xf = x.detach().float().contiguous()
            t = timesteps.to(device=x.device, dtype=torch.long).contiguous()
            ctx = context.detach().to(device=x.device).float().contiguous()
            assert ctx.shape[0] == xf.shape[0] and ctx.shape[2] == self.context_dim
            if c_input is not None or getattr(self, "_cin_active", False):
                self.set_c_input(None if c_input is None else c_input.to(x.device), xf.shape[3])
                self._cin_active = c_input is not None
            out = self.forward_native(xf, t, ctx)
        if torch.is_autocast_enabled():
            return out.half()  # what the reference returns under torch.autocast("cuda")
        return out.to(x.dtype) if x.dtype != torch.float32 else out

    def last_flops(self):
        return N.lib().lr_unet_last_flops(self.engine())


class MultiViewUnetModel(UNetModel):
    """ldm/modules/diffusionmodules/multiview_unet.py:33-411 — same UNet, self-attention runs across `view_num`
    views of a sample (multiview_attention.py:431-468)."""

    def __init__(self, *args, view_num=4, concat_target=False, **kwargs):
        super().__init__(*args, view_num=view_num, concat_target=concat_target, **kwargs)


class NVSUnetModel(UNetModel):
    """inpainting_ldm/NVS_ldm.py:22-104. `use_sep=True` adds the learned `sep_token.<channels>` parameters and the engine
    inserts / removes the separator column around every non-resampling block (:57-97, odd feature widths W + 1);
    `forward(..., c_input=...)` adds the refinement features to the input conv's output (:49,64-68). With
    `use_sep=False` and no c_input (configs/novel_view_synthesis.yaml) it is the plain UNetModel."""

    def __init__(self, *args, **kwargs):
        kwargs.setdefault("use_sep", False)
        super().__init__(*args, **kwargs)


class _EngineHandle:
    def __init__(self, h):
        self.h = h

    def __del__(self):
        try:
            N.lib().lr_unet_destroy(self.h)
        except Exception:
            pass
