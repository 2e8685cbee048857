"""Drop-in transformer modules (reference ldm/modules/attention.py): GEGLU :51-58, FeedForward :61-78,
CrossAttention :147-196, BasicTransformerBlock :253-283, SpatialTransformer :331-419.

Inside `UNetModel.forward` these classes are parameter containers only (the engine runs the whole graph natively).
Their own `forward` methods exist for stand-alone use and call the same sm_100a kernels through the op-level C ABI:
tcgen05 GEMMs with fused bias/residual/GEGLU epilogues, the fused softmax(QK^T)V kernel and the LayerNorm/GroupNorm
kernels. No PyTorch math on the hot path, no fallback.
"""
import torch
import torch.nn as nn

from . import ops


def zero_module(module):
    for p in module.parameters():
        p.detach().zero_()
    return module


def Normalize(in_channels):
    return nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)


def _tokens_f16(x):
    assert x.is_cuda, "leftrefill_b200 modules run on CUDA only"
    return x.detach().to(torch.float16).contiguous()


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        shp = x.shape
        w = ops.repack_linear(self.proj.weight, geglu=True)
        n = self.proj.out_features // 2
        b = torch.empty_like(self.proj.bias, dtype=torch.float32)
        b[0::2], b[1::2] = self.proj.bias[:n].float(), self.proj.bias[n:].float()
        y = ops.linear(_tokens_f16(x).reshape(-1, shp[-1]), w, bias=b, geglu=True)
        return y.reshape(*shp[:-1], n).to(x.dtype)


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, glu=False, dropout=0.):
        super().__init__()
        if not glu:
            raise NotImplementedError("only the gated (GEGLU) feed-forward used by BasicTransformerBlock is implemented")
        inner = int(dim * mult)
        self.net = nn.Sequential(GEGLU(dim, inner), nn.Dropout(dropout), nn.Linear(inner, dim_out or dim))

    def forward(self, x):
        h = self.net[0](x)
        lin = self.net[2]
        y = ops.linear(_tokens_f16(h).reshape(-1, h.shape[-1]), ops.repack_linear(lin.weight), bias=lin.bias.float())
        return y.reshape(*x.shape[:-1], lin.out_features).to(x.dtype)


class CrossAttention(nn.Module):
    """softmax(q k^T / sqrt(d)) v with d_head = 64; self-attention when `context` is None."""

    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64, dropout=0.):
        super().__init__()
        inner = dim_head * heads
        context_dim = context_dim if context_dim is not None else query_dim
        self.scale, self.heads, self.dim_head = dim_head ** -0.5, heads, dim_head
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(context_dim, inner, bias=False)
        self.to_v = nn.Linear(context_dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, query_dim), nn.Dropout(dropout))

    def forward(self, x, context=None, mask=None):
        if mask is not None:
            raise NotImplementedError("attention masks are never passed on the LeftRefill path (attention.py:292)")
        if self.dim_head != 64:
            raise NotImplementedError("the fused attention kernel supports dim_head == 64 only")
        b, n, _ = x.shape
        xh = _tokens_f16(x)
        ch = xh if context is None else _tokens_f16(context)
        m = ch.shape[1]
        inner = self.heads * self.dim_head
        q = ops.linear(xh.reshape(b * n, -1), ops.repack_linear(self.to_q.weight)).reshape(b, n, inner)
        k = ops.linear(ch.reshape(b * m, -1), ops.repack_linear(self.to_k.weight)).reshape(b, m, inner)
        v = ops.linear(ch.reshape(b * m, -1), ops.repack_linear(self.to_v.weight)).reshape(b, m, inner)
        o = ops.attention(q, k, v, self.heads, self.scale)
        lin = self.to_out[0]
        y = ops.linear(o.reshape(b * n, inner), ops.repack_linear(lin.weight), bias=lin.bias.float())
        return y.reshape(b, n, -1).to(x.dtype)


MemoryEfficientCrossAttention = CrossAttention  # the xformers variant has the same parameters and semantics


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, n_heads, d_head, dropout=0., context_dim=None, gated_ff=True, checkpoint=True,
                 disable_self_attn=False, **kwargs):
        super().__init__()
        if disable_self_attn:
            raise NotImplementedError("disable_self_attn is not used by any LeftRefill config")
        self.disable_self_attn = False
        # registration order fixes the state-dict key order: attn1, ff, attn2, norm1..3 (attention.py:263-270)
        self.attn1 = CrossAttention(query_dim=dim, heads=n_heads, dim_head=d_head, dropout=dropout)
        self.ff = FeedForward(dim, dropout=dropout, glu=gated_ff)
        self.attn2 = CrossAttention(query_dim=dim, context_dim=context_dim, heads=n_heads, dim_head=d_head,
                                    dropout=dropout)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.checkpoint = checkpoint

    def _ln(self, norm, x):
        return ops.layernorm(_tokens_f16(x), norm.weight.float(), norm.bias.float(), norm.eps)

    def forward(self, x, context=None):
        x = self.attn1(self._ln(self.norm1, x)) + x.to(torch.float16)
        x = self.attn2(self._ln(self.norm2, x), context=context) + x
        x = self.ff(self._ln(self.norm3, x)) + x
        return x


class SpatialTransformer(nn.Module):
    def __init__(self, in_channels, n_heads, d_head, depth=1, dropout=0., context_dim=None, disable_self_attn=False,
                 use_linear=False, use_checkpoint=True, one_attn=False, num_patches=None):
        super().__init__()
        if one_attn or disable_self_attn:
            raise NotImplementedError("one_attn / disable_self_attn are not used by any LeftRefill config")
        if context_dim is not None and not isinstance(context_dim, list):
            context_dim = [context_dim] * depth
        self.in_channels = in_channels
        inner = n_heads * d_head
        self.norm = Normalize(in_channels)
        self.proj_in = nn.Linear(in_channels, inner) if use_linear else nn.Conv2d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner, n_heads, d_head, dropout=dropout, context_dim=context_dim[d],
                                  checkpoint=use_checkpoint) for d in range(depth)])
        self.proj_out = zero_module(nn.Linear(in_channels, inner) if use_linear else nn.Conv2d(inner, in_channels, 1))
        self.use_linear = use_linear

    def forward(self, x, context=None, **kwargs):
        if not isinstance(context, list):
            context = [context]
        b, c, h, w = x.shape
        t = ops.to_nhwc_f16(x)  # [b, h, w, c]
        xn = ops.groupnorm(t, self.norm.weight.float(), self.norm.bias.float(), self.norm.eps, silu=False)
        wi = ops.repack_linear(self.proj_in.weight.reshape(self.proj_in.weight.shape[0], -1))
        tok = ops.linear(xn.reshape(b * h * w, c), wi, bias=self.proj_in.bias.float()).reshape(b, h * w, -1)
        for i, blk in enumerate(self.transformer_blocks):
            tok = blk(tok, context=context[i] if i < len(context) else context[-1])
        wo = ops.repack_linear(self.proj_out.weight.reshape(self.proj_out.weight.shape[0], -1))
        y = ops.linear(tok.reshape(b * h * w, -1), wo, bias=self.proj_out.bias.float(), residual=t.reshape(b * h * w, c))
        y = ops.to_nchw_f32(y.reshape(b, h, w, c))
        return y.to(x.dtype) if x.dtype != torch.float32 else y
