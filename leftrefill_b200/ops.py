"""Op-level Python wrappers over the C ABI (torch tensors are only device-memory handles here).

Activations are NHWC fp16 on the GPU: a feature map is [n, h, w, c] (equivalently a token matrix [n*h*w, c]).
"""
import torch

from . import _native as N


def _chk(t, dtype=None):
    assert t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
    if dtype is not None:
        assert t.dtype == dtype, f"expected {dtype}, got {t.dtype}"
    return t


def repack_conv3x3(w_oihw):
    """fp32 [O, I, 3, 3] -> fp16 [O, 9*I] with k = (ky*3+kx)*I + i."""
    w = _chk(w_oihw.float().contiguous())
    O, I = w.shape[0], w.shape[1]
    out = torch.empty(O, 9 * I, dtype=torch.float16, device=w.device)
    N.check(N.lib().lr_repack_conv3x3_weight(N.ptr(w), O, I, N.ptr(out), N.current_stream()), "repack_conv3x3")
    return out


def geglu_interleave(t):
    """Row order of a fused GEGLU projection (weights [2n, I] or bias [2n]): source rows [0, n) values, [n, 2n) gates ->
    groups of four (value_2k, value_2k+1, gate_2k, gate_2k+1), the order lr_repack_linear_weight(geglu=1) produces."""
    n = t.shape[0] // 2
    v, g = t[:n], t[n:]
    return torch.stack([v[0::2], v[1::2], g[0::2], g[1::2]], dim=1).reshape(t.shape).contiguous()


def repack_linear(w, geglu=False):
    """fp32 [O, I] -> fp16 [O, I]; geglu puts the rows in the order of geglu_interleave."""
    w = _chk(w.float().contiguous())
    O, I = w.shape
    out = torch.empty(O, I, dtype=torch.float16, device=w.device)
    N.check(N.lib().lr_repack_linear_weight(N.ptr(w), O, I, int(geglu), N.ptr(out), N.current_stream()),
            "repack_linear")
    return out


def linear(a, w, bias=None, residual=None, geglu=False, force_block_n=0):
    """a fp16 [M, K]; w fp16 [n_w, K] (n_w = 2*n_out when geglu). Returns fp16 [M, n_out]."""
    _chk(a, torch.float16)
    _chk(w, torch.float16)
    M, K = a.shape
    n_w = w.shape[0]
    n_out = n_w // 2 if geglu else n_w
    out = torch.empty(M, n_out, dtype=torch.float16, device=a.device)
    if bias is not None:
        _chk(bias, torch.float32)
    if residual is not None:
        _chk(residual, torch.float16)
    N.check(N.lib().lr_linear_f16(N.ptr(a), a.stride(0), M, K, N.ptr(w), w.stride(0), n_out, N.ptr(bias),
                                  N.ptr(residual), residual.stride(0) if residual is not None else 0, N.ptr(out),
                                  out.stride(0), int(geglu), force_block_n, N.current_stream()), "linear")
    return out


def conv3x3(x0, wt, bias=None, x1=None, stride=1, bias_img=None, residual=None, force_block_n=0):
    """x0 fp16 NHWC [n, h, w, c0] (+ optional x1 [n, h, w, c1]); wt fp16 [cout, 9*(c0+c1)]. Returns [n, ho, wo, cout]."""
    _chk(x0, torch.float16)
    n, h, w, c0 = x0.shape
    c1 = 0
    if x1 is not None:
        _chk(x1, torch.float16)
        c1 = x1.shape[3]
    cout = wt.shape[0]
    ho, wo = (h, w) if stride == 1 else ((h - 1) // 2 + 1, (w - 1) // 2 + 1)
    out = torch.empty(n, ho, wo, cout, dtype=torch.float16, device=x0.device)
    N.check(N.lib().lr_conv3x3_f16(N.ptr(x0), c0, N.ptr(x1), c1, n, h, w, stride, N.ptr(_chk(wt, torch.float16)), cout,
                                   N.ptr(bias), N.ptr(bias_img), N.ptr(residual), N.ptr(out), force_block_n,
                                   N.current_stream()), "conv3x3")
    return out


def upsample2x_conv3x3(x, wt, bias=None):
    """nearest x2 upsample + 3x3 conv (pad 1) of NHWC fp16 x [n, h, w, cin] without the upsampled tensor: four folded
    2x2-tap phase convs. wt fp16 [cout, 9*cin] (repack_conv3x3). Returns [n, 2h, 2w, cout]."""
    _chk(x, torch.float16)
    n, h, w, cin = x.shape
    cout = wt.shape[0]
    out = torch.empty(n, 2 * h, 2 * w, cout, dtype=torch.float16, device=x.device)
    scratch = torch.empty(16 * cout * cin, dtype=torch.float16, device=x.device)
    N.check(N.lib().lr_upsample2x_conv3x3_f16(N.ptr(x), n, h, w, cin, N.ptr(_chk(wt, torch.float16)), cout, N.ptr(bias),
                                              N.ptr(scratch), N.ptr(out), N.current_stream()), "upsample2x_conv3x3")
    return out


def attention(q, k, v, heads, scale=None):
    """q [b, tq, heads*64], k/v [b, tk, heads*64] fp16 (may be column slices of wider tensors). Returns [b, tq, heads*64]."""
    assert q.dtype == k.dtype == v.dtype == torch.float16
    b, tq, c = q.shape
    tk = k.shape[1]
    assert c == heads * 64, "d_head must be 64"
    if scale is None:
        scale = 64 ** -0.5
    for t in (q, k, v):
        assert t.is_cuda and t.stride(2) == 1 and t.stride(0) == t.shape[1] * t.stride(1)
    out = torch.empty(b, tq, c, dtype=torch.float16, device=q.device)
    N.check(N.lib().lr_attention_f16(N.ptr(q), q.stride(1), 0, N.ptr(k), k.stride(1), 0, N.ptr(v), v.stride(1), 0,
                                     N.ptr(out), c, b, heads, tq, tk, float(scale), N.current_stream()), "attention")
    return out


def groupnorm(x0, gamma, beta, eps, silu=False, x1=None, groups=32):
    """GroupNorm(+SiLU) over the channel concat of NHWC fp16 x0 (and x1). Returns fp16 [n, h, w, c0+c1]."""
    _chk(x0, torch.float16)
    n, h, w, c0 = x0.shape
    c1 = x1.shape[3] if x1 is not None else 0
    C = c0 + c1
    out = torch.empty(n, h, w, C, dtype=torch.float16, device=x0.device)
    scratch = torch.empty(N.lib().lr_groupnorm_scratch_bytes(n, groups, h * w), dtype=torch.uint8, device=x0.device)
    N.check(N.lib().lr_groupnorm_f16(N.ptr(x0), c0, N.ptr(x1), c1, n, h * w, groups, float(eps),
                                     N.ptr(_chk(gamma, torch.float32)), N.ptr(_chk(beta, torch.float32)), int(silu),
                                     N.ptr(out), N.ptr(scratch), N.current_stream()), "groupnorm")
    return out


class GNStats:
    """Per-tile (sum, sum of squares) partials a conv / linear left for its OUTPUT (lr_gn_conv3x3_f16 / lr_gn_linear_f16):
    table fp32 [rows, C, 2], `ppi` rows per image, `P` pixels per image. Feeds gn_finalize."""

    def __init__(self, table, ppi, C, P):
        self.table, self.ppi, self.C, self.P = table, ppi, C, P


def gn_finalize(stats0, gamma, beta, eps, n, stats1=None, groups=32):
    """GroupNorm statistics from producer partials -> (scale, shift) fp32 [n, c0 + c1] with
    y = x * scale + shift == GroupNorm(x) (elementwise.cuh gn_finalize_kernel)."""
    C = stats0.C + (stats1.C if stats1 is not None else 0)
    scale = torch.empty(n, C, dtype=torch.float32, device=stats0.table.device)
    shift = torch.empty_like(scale)
    N.check(N.lib().lr_gn_finalize(N.ptr(stats0.table), stats0.ppi, stats0.C,
                                   N.ptr(stats1.table) if stats1 is not None else None,
                                   stats1.ppi if stats1 is not None else 0, stats1.C if stats1 is not None else 0, n,
                                   stats0.P, groups, float(eps), N.ptr(_chk(gamma, torch.float32)),
                                   N.ptr(_chk(beta, torch.float32)), N.ptr(scale), N.ptr(shift), N.current_stream()),
            "gn_finalize")
    return scale, shift


def gn_conv3x3(x0, wt, gn=None, silu=True, bias=None, x1=None, bias_img=None, residual=None, want_stats=False,
               force_block_n=0):
    """3x3 conv (stride 1, pad 1) whose input is GroupNorm(+SiLU) of concat(x0, x1) given as gn = (scale, shift) fp32
    [n, c0 + c1] - applied to the activation tiles in shared memory - and/or whose epilogue leaves GroupNorm partials
    of its output. Returns out, or (out, GNStats | None) when want_stats."""
    import ctypes
    _chk(x0, torch.float16)
    n, h, w, c0 = x0.shape
    c1 = x1.shape[3] if x1 is not None else 0
    cout = wt.shape[0]
    out = torch.empty(n, h, w, cout, dtype=torch.float16, device=x0.device)
    table = None
    if want_stats:
        rows = N.lib().lr_conv_stats_rows(n, h, w, 1, 9, 0)
        if rows > 0:
            table = torch.zeros(rows, cout, 2, dtype=torch.float32, device=x0.device)
    ppi = ctypes.c_int(0)
    sc, sh = gn if gn is not None else (None, None)
    N.check(N.lib().lr_gn_conv3x3_f16(N.ptr(x0), c0, N.ptr(x1), c1, n, h, w, N.ptr(sc), N.ptr(sh), int(silu),
                                      N.ptr(_chk(wt, torch.float16)), cout, N.ptr(bias), N.ptr(bias_img), N.ptr(residual),
                                      N.ptr(out), N.ptr(table), ctypes.byref(ppi), force_block_n, N.current_stream()),
            "gn_conv3x3")
    if not want_stats:
        return out
    return out, (GNStats(table, ppi.value, cout, h * w) if ppi.value > 0 else None)


def gn_linear(a, w, rows_per_img, gn=None, silu=False, bias=None, residual=None, want_stats=False, force_block_n=0):
    """Linear on a token matrix a [M, K] (image of row m = m // rows_per_img) with the same two fusions as gn_conv3x3."""
    import ctypes
    _chk(a, torch.float16)
    M, K = a.shape
    n_out = w.shape[0]
    out = torch.empty(M, n_out, dtype=torch.float16, device=a.device)
    table = None
    if want_stats:
        rows = N.lib().lr_conv_stats_rows(1, 1, M, 1, 1, rows_per_img)
        if rows > 0:
            table = torch.zeros(rows, n_out, 2, dtype=torch.float32, device=a.device)
    ppi = ctypes.c_int(0)
    sc, sh = gn if gn is not None else (None, None)
    N.check(N.lib().lr_gn_linear_f16(N.ptr(a), M, K, rows_per_img, N.ptr(sc), N.ptr(sh), int(silu),
                                     N.ptr(_chk(w, torch.float16)), n_out, N.ptr(bias), N.ptr(residual), N.ptr(out),
                                     N.ptr(table), ctypes.byref(ppi), force_block_n, N.current_stream()), "gn_linear")
    if not want_stats:
        return out
    return out, (GNStats(table, ppi.value, n_out, rows_per_img) if ppi.value > 0 else None)


def layernorm(x, gamma, beta, eps=1e-5):
    _chk(x, torch.float16)
    C = x.shape[-1]
    M = x.numel() // C
    out = torch.empty_like(x)
    N.check(N.lib().lr_layernorm_f16(N.ptr(x), M, C, N.ptr(_chk(gamma, torch.float32)),
                                     N.ptr(_chk(beta, torch.float32)), float(eps), N.ptr(out), N.current_stream()),
            "layernorm")
    return out


def to_nhwc_f16(x_nchw):
    x = _chk(x_nchw.float().contiguous())
    n, c, h, w = x.shape
    out = torch.empty(n, h, w, c, dtype=torch.float16, device=x.device)
    N.check(N.lib().lr_nchw_f32_to_nhwc_f16(N.ptr(x), n, c, h, w, N.ptr(out), N.current_stream()), "to_nhwc")
    return out


def to_nchw_f32(x_nhwc):
    _chk(x_nhwc, torch.float16)
    n, h, w, c = x_nhwc.shape
    out = torch.empty(n, c, h, w, dtype=torch.float32, device=x_nhwc.device)
    N.check(N.lib().lr_nhwc_f16_to_nchw_f32(N.ptr(x_nhwc), c, n, c, h, w, N.ptr(out), N.current_stream()), "to_nchw")
    return out


def ddim_update(x, eps_uncond, eps_cond, noise, cfg_scale, a_t, a_prev, sigma_t, sqrt_one_minus_at, temperature=1.0):
    """Fused CFG combine + DDIM update (ldm/models/diffusion/ddim.py:343,359-381). Returns (x_prev, pred_x0)."""
    _chk(x, torch.float32)
    x_prev = torch.empty_like(x)
    pred_x0 = torch.empty_like(x)
    N.check(N.lib().lr_ddim_update(N.ptr(x), N.ptr(_chk(eps_uncond, torch.float32)),
                                   N.ptr(eps_cond), N.ptr(noise), float(cfg_scale), float(a_t), float(a_prev),
                                   float(sigma_t), float(sqrt_one_minus_at), float(temperature), x.numel(),
                                   N.ptr(x_prev), N.ptr(pred_x0), N.current_stream()), "ddim_update")
    return x_prev, pred_x0


def ddim_update_dev(x, eps_uncond, eps_cond, noise, coef, temperature, x_prev, pred_x0):
    """ddim_update with the step scalars in device memory: coef fp32 [5] = (cfg, a_t, a_prev, sigma_t, sqrt(1 - a_t)).
    Writes into the given x_prev (may alias x) / pred_x0 buffers; CUDA-graph capturable (no allocation)."""
    _chk(x, torch.float32)
    _chk(coef, torch.float32)
    assert coef.numel() >= 5
    N.check(N.lib().lr_ddim_update_dev(N.ptr(x), N.ptr(_chk(eps_uncond, torch.float32)), N.ptr(eps_cond),
                                       N.ptr(noise), N.ptr(coef), float(temperature), x.numel(), N.ptr(_chk(x_prev)),
                                       N.ptr(_chk(pred_x0)), N.current_stream()), "ddim_update_dev")
    return x_prev, pred_x0
