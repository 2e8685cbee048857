"""GPU bring-up diagnostics for the op-level C ABI. Each case runs in its own subprocess (a trapped kernel poisons the
CUDA context) with a timeout. Usage on the GPU box:

    python tests/gpu_diag_ops.py            # run all cases, print a table
    python tests/gpu_diag_ops.py --case X   # run one case in-process
"""
import argparse
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def report(name, got, ref, tol=None, rtol=1e-3, atol_scale=2e-4):
    """Element-wise criterion |got - ref| <= atol + rtol*|ref| with rtol = 1e-3 (BASELINE north_star) and
    atol = 2e-4 * max(1, max|ref|): one fp16 rounding of the output (2^-11 relative) plus fp32 accumulation noise."""
    import torch
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    scale = ref.abs().max().item() + 1e-9
    bound = atol_scale * max(1.0, scale) + rtol * ref.abs()
    bad = err > bound
    msg = (f"[{name}] max_abs_err={err.max().item():.4e} ref_max={scale:.4e} worst_ratio={(err / bound).max().item():.3f} "
           f"bad_frac={bad.float().mean().item():.5f} finite={bool(torch.isfinite(got).all())}")
    if bad.any():
        idx = bad.nonzero()[:6].tolist()
        msg += f" first_bad={idx}"
        if got.dim() == 2:
            rows = bad.any(dim=1).nonzero().flatten()
            cols = bad.any(dim=0).nonzero().flatten()
            msg += (f" bad_rows={rows[:12].tolist()}(+{max(0, rows.numel() - 12)})"
                    f" bad_cols={cols[:12].tolist()}(+{max(0, cols.numel() - 12)})")
    ok = (not bad.any().item()) and bool(torch.isfinite(got).all())
    print(("PASS " if ok else "FAIL ") + msg, flush=True)
    return ok


def case_linear(M, K, Nn, bias=False, residual=False, geglu=False, force=0, seed=0):
    import torch
    from leftrefill_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(seed)
    a = (torch.randn(M, K, generator=g) * 0.5).cuda().half()
    n_w = 2 * Nn if geglu else Nn
    w32 = (torch.randn(n_w, K, generator=g) / K ** 0.5).cuda()
    b = torch.randn(n_w, generator=g).cuda() if bias else None
    r = torch.randn(M, Nn, generator=g).cuda().half() if residual else None
    wp = ops.repack_linear(w32, geglu=geglu)
    bp = b
    if geglu and b is not None:
        bp = ops.geglu_interleave(b)
    out = ops.linear(a, wp, bias=bp, residual=r, geglu=geglu, force_block_n=force)
    torch.cuda.synchronize()
    ref = a.float() @ w32.half().float().t()
    if b is not None:
        ref = ref + b
    if geglu:
        ref = ref[:, :Nn] * torch.nn.functional.gelu(ref[:, Nn:])
    if r is not None:
        ref = ref + r.float()
    return report(f"linear M={M} K={K} N={Nn} bias={bias} res={residual} geglu={geglu} bn={force}", out, ref)


def case_conv(n, h, w, c0, cout, c1=0, stride=1, bias_img=False, residual=False, force=0, seed=0):
    import torch
    import torch.nn.functional as F
    from leftrefill_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = (torch.randn(n, c0 + c1, h, w, generator=g)).cuda()
    wt = (torch.randn(cout, c0 + c1, 3, 3, generator=g) / (9 * (c0 + c1)) ** 0.5).cuda()
    b = torch.randn(cout, generator=g).cuda()
    x0 = ops.to_nhwc_f16(x[:, :c0].contiguous())
    x1 = ops.to_nhwc_f16(x[:, c0:].contiguous()) if c1 else None
    bi = torch.randn(n, cout, generator=g).cuda() if bias_img else None
    ho, wo = (h, w) if stride == 1 else ((h - 1) // 2 + 1, (w - 1) // 2 + 1)
    r = torch.randn(n, ho, wo, cout, generator=g).cuda().half() if residual else None
    wp = ops.repack_conv3x3(wt)
    out = ops.conv3x3(x0, wp, bias=b, x1=x1, stride=stride, bias_img=bi, residual=r, force_block_n=force)
    torch.cuda.synchronize()
    ref = F.conv2d(x.half().float(), wt.half().float(), b, stride=stride, padding=1)
    if bi is not None:
        ref = ref + bi[:, :, None, None]
    ref = ref.permute(0, 2, 3, 1)
    if r is not None:
        ref = ref + r.float()
    return report(f"conv n={n} {h}x{w} c={c0}+{c1}->{cout} s={stride} bimg={bias_img} res={residual} bn={force}",
                  out.reshape(-1, cout), ref.reshape(-1, cout))


def case_upconv(n, h, w, cin, cout, seed=0):
    """Upsample (nearest x2 + conv3x3) folded into four 2x2-tap phase convs vs F.interpolate + F.conv2d."""
    import torch
    import torch.nn.functional as F
    from leftrefill_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(n, cin, h, w, generator=g).cuda()
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5).cuda()
    b = torch.randn(cout, generator=g).cuda()
    x0 = ops.to_nhwc_f16(x)
    out = ops.upsample2x_conv3x3(x0, ops.repack_conv3x3(wt), bias=b)
    torch.cuda.synchronize()
    ref = F.conv2d(F.interpolate(x.half().float(), scale_factor=2, mode="nearest"), wt.half().float(), b, padding=1)
    # the folded weights are fp16(sum of fp16 taps): one extra weight rounding (2^-11) -> rtol 2e-3
    return report(f"upsample+conv n={n} {h}x{w} {cin}->{cout}", out.reshape(-1, cout),
                  ref.permute(0, 2, 3, 1).reshape(-1, cout), rtol=2e-3, atol_scale=4e-4)


def case_attn(b, heads, tq, tk, fused_qkv=False, seed=0):
    import torch
    from leftrefill_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(seed)
    C = heads * 64
    if fused_qkv:
        assert tq == tk
        qkv = torch.randn(b, tq, 3 * C, generator=g).cuda().half()
        q, k, v = qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:]
    else:
        q = torch.randn(b, tq, C, generator=g).cuda().half()
        k = torch.randn(b, tk, C, generator=g).cuda().half()
        v = torch.randn(b, tk, C, generator=g).cuda().half()
    out = ops.attention(q, k, v, heads)
    torch.cuda.synchronize()
    qf = q.float().reshape(b, tq, heads, 64).permute(0, 2, 1, 3)
    kf = k.float().reshape(b, tk, heads, 64).permute(0, 2, 1, 3)
    vf = v.float().reshape(b, tk, heads, 64).permute(0, 2, 1, 3)
    sim = (qf @ kf.transpose(-1, -2)) * 0.125
    ref = (sim.softmax(-1) @ vf).permute(0, 2, 1, 3).reshape(b, tq, C)
    return report(f"attn b={b} h={heads} tq={tq} tk={tk} fused={fused_qkv}", out.reshape(-1, C), ref.reshape(-1, C))


def case_attn_bigrange(b, heads, tq, tk, logit_std=16.0, ramp=40.0, seed=0):
    """Adversarial range for the online softmax (attention_tc.cuh lazy rescale, threshold 2^8): scaled logits with
    standard deviation `logit_std` plus a ramp of `ramp` (log units) along the key index, so that the running row maximum
    keeps increasing from K/V tile to K/V tile and the TMEM-resident O is rescaled many times."""
    import torch
    from leftrefill_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(seed)
    C = heads * 64
    s = logit_std ** 0.5                       # q.k/8 over 64 unit-variance products has std s^2
    q = torch.randn(b, tq, C, generator=g) * s
    k = torch.randn(b, tk, C, generator=g) * s
    v = torch.randn(b, tk, C, generator=g)
    u = torch.randn(heads, 64, generator=g)
    u = u / u.norm(dim=1, keepdim=True)
    # q gets +a*u, k_j gets +(j/tk)*c*u: the logit gains a*c*(j/tk)/8 -> choose a*c/8 = ramp
    a = 4.0
    c = ramp * 8.0 / a
    q = q + a * u.reshape(1, 1, C)
    k = k + (torch.arange(tk).float() / tk).reshape(1, tk, 1) * c * u.reshape(1, 1, C)
    q, k, v = q.cuda().half(), k.cuda().half(), v.cuda().half()
    out = ops.attention(q, k, v, heads)
    torch.cuda.synchronize()
    qf = q.double().reshape(b, tq, heads, 64).permute(0, 2, 1, 3)
    kf = k.double().reshape(b, tk, heads, 64).permute(0, 2, 1, 3)
    vf = v.double().reshape(b, tk, heads, 64).permute(0, 2, 1, 3)
    sim = (qf @ kf.transpose(-1, -2)) * 0.125
    ref = (sim.softmax(-1) @ vf).permute(0, 2, 1, 3).reshape(b, tq, C).float()
    return report(f"attn bigrange b={b} h={heads} tq={tq} tk={tk} std={logit_std} ramp={ramp}", out.reshape(-1, C),
                  ref.reshape(-1, C))


def case_gn_bigmean(n, h, w, c, ratio=50.0, silu=True, seed=0):
    """GroupNorm with |mean| / std >= `ratio` in every group (real SD2 checkpoints have such outlier channels): the
    E[x^2] - mean^2 form cancels ~ratio^2 of its leading digits."""
    import torch
    import torch.nn.functional as F
    from leftrefill_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(seed)
    sign = torch.where(torch.arange(c) % 64 < 32, 1.0, -1.0).reshape(1, c, 1, 1)   # groups alternate in sign
    x = (torch.randn(n, c, h, w, generator=g) * 0.25 + 0.25 * ratio * sign).cuda()
    gamma = torch.randn(c, generator=g).cuda()
    beta = torch.randn(c, generator=g).cuda()
    x0 = ops.to_nhwc_f16(x)
    out = ops.groupnorm(x0, gamma, beta, 1e-5, silu=silu)
    torch.cuda.synchronize()
    xr = x0.float().permute(0, 3, 1, 2).double()
    ref = F.group_norm(xr, 32, gamma.double(), beta.double(), 1e-5)
    if silu:
        ref = F.silu(ref)
    return report(f"groupnorm bigmean n={n} {h}x{w} c={c} |mean|/std={ratio}", out.reshape(-1, c),
                  ref.float().permute(0, 2, 3, 1).reshape(-1, c))


def case_geglu_big(M, K, Nn, seed=0):
    """GEGLU with outputs of the order of 3e4 (fp16 max 65504): value ~ +-150, gate ~ +-150 -> v * gelu(g) up to ~4e4;
    the product must be formed in fp32 and only the result rounded to fp16."""
    import torch
    from leftrefill_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(seed)
    a = (torch.randn(M, K, generator=g)).cuda().half()
    w32 = (torch.randn(2 * Nn, K, generator=g) * (50.0 / K ** 0.5)).cuda()
    b = (torch.randn(2 * Nn, generator=g) * 20.0).cuda()
    wp = ops.repack_linear(w32, geglu=True)
    bp = ops.geglu_interleave(b)
    out = ops.linear(a, wp, bias=bp, geglu=True)
    torch.cuda.synchronize()
    h = a.double() @ w32.half().double().t() + b.double()
    ref = (h[:, :Nn] * torch.nn.functional.gelu(h[:, Nn:])).float()
    ok_range = ref.abs().max().item() > 2.0e4 and ref.abs().max().item() < 6.5e4
    print(f"  geglu_big ref range max|ref|={ref.abs().max().item():.1f} (want 2e4..6.5e4: {ok_range})")
    return report(f"geglu big M={M} K={K} N={Nn}", out, ref) and ok_range


def case_gn(n, h, w, c0, c1=0, silu=True, eps=1e-5, seed=0):
    import torch
    import torch.nn.functional as F
    from leftrefill_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(seed)
    C = c0 + c1
    x = (torch.randn(n, C, h, w, generator=g) * 2 + 0.5).cuda()
    gamma = torch.randn(C, generator=g).cuda()
    beta = torch.randn(C, generator=g).cuda()
    x0 = ops.to_nhwc_f16(x[:, :c0].contiguous())
    x1 = ops.to_nhwc_f16(x[:, c0:].contiguous()) if c1 else None
    out = ops.groupnorm(x0, gamma, beta, eps, silu=silu, x1=x1)
    torch.cuda.synchronize()
    ref = F.group_norm(x.half().float(), 32, gamma, beta, eps)
    if silu:
        ref = F.silu(ref)
    return report(f"groupnorm n={n} {h}x{w} c={c0}+{c1} silu={silu}", out.reshape(-1, C),
                  ref.permute(0, 2, 3, 1).reshape(-1, C))


def case_gn_fused_conv(n, h, w, c0, cout, c1=0, silu=True, eps=1e-5, residual=False, force=0, bigmean=False, seed=0):
    """The fused GroupNorm path end to end at op level: a producer conv leaves statistics partials for its output(s),
    lr_gn_finalize turns them into (scale, shift), and the consumer conv applies GroupNorm + SiLU to its activation tiles
    in shared memory. Reference: F.group_norm + F.silu + F.conv2d on the producers' fp16 outputs."""
    import torch
    import torch.nn.functional as F
    from leftrefill_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cpu").manual_seed(seed)
    C = c0 + c1

    def producer(cin, co):
        x = torch.randn(n, cin, h, w, generator=g).cuda()
        wt = (torch.randn(co, cin, 3, 3, generator=g) / (9 * cin) ** 0.5).cuda()
        b = torch.randn(co, generator=g).cuda() * (6.0 if bigmean else 1.0)
        y, st = ops.gn_conv3x3(ops.to_nhwc_f16(x), ops.repack_conv3x3(wt), bias=b, want_stats=True)
        return y, st

    y0, st0 = producer(64, c0)
    y1, st1 = producer(64, c1) if c1 else (None, None)
    assert st0 is not None and (c1 == 0 or st1 is not None), "producer did not leave statistics"
    gamma = torch.randn(C, generator=g).cuda()
    beta = torch.randn(C, generator=g).cuda()
    scale, shift = ops.gn_finalize(st0, gamma, beta, eps, n, stats1=st1)
    xcat = torch.cat([y0, y1], dim=3) if c1 else y0
    xf = xcat.float().permute(0, 3, 1, 2).double()
    # statistics check: scale / shift against fp64 group statistics of the fp16 producer outputs
    xg = xf.reshape(n, 32, C // 32, h, w)
    mean = xg.mean(dim=(2, 3, 4))
    rstd = 1.0 / (xg.var(dim=(2, 3, 4), unbiased=False) + eps).sqrt()
    sc_ref = (rstd.repeat_interleave(C // 32, dim=1) * gamma.double()[None]).float()
    sh_ref = (beta.double()[None] - mean.repeat_interleave(C // 32, dim=1) * sc_ref.double()).float()
    ok = report(f"gn-finalize scale n={n} {h}x{w} c={c0}+{c1}", scale, sc_ref, rtol=2e-4 if bigmean else 2e-5, atol_scale=1e-6)
    ok &= report(f"gn-finalize shift n={n} {h}x{w} c={c0}+{c1}", shift, sh_ref, rtol=2e-4 if bigmean else 2e-5,
                 atol_scale=2e-4 if bigmean else 2e-6)
    wt = (torch.randn(cout, C, 3, 3, generator=g) / (9 * C) ** 0.5).cuda()
    b = torch.randn(cout, generator=g).cuda()
    r = torch.randn(n, h, w, cout, generator=g).cuda().half() if residual else None
    out, st_out = ops.gn_conv3x3(y0, ops.repack_conv3x3(wt), gn=(scale, shift), silu=silu, bias=b, x1=y1, residual=r,
                                 want_stats=True, force_block_n=force)
    torch.cuda.synchronize()
    # reference with the SAME coefficients (isolates the transform + conv) ...
    xn = xf.float() * scale[:, :, None, None] + shift[:, :, None, None]
    if silu:
        xn = F.silu(xn)
    xn = xn.half().float()                       # the transform rounds to fp16 before the MMA
    ref = F.conv2d(xn, wt.half().float(), b, padding=1).permute(0, 2, 3, 1)
    if r is not None:
        ref = ref + r.float()
    ok &= report(f"gn-fused conv n={n} {h}x{w} c={c0}+{c1}->{cout} silu={silu} res={residual} bn={force}",
                 out.reshape(-1, cout), ref.reshape(-1, cout))
    # ... and the statistics the consumer left for ITS output
    if st_out is not None:
        s2, h2 = ops.gn_finalize(st_out, torch.ones(cout, device="cuda"), torch.zeros(cout, device="cuda"), eps, n)
        og = out.float().permute(0, 3, 1, 2).double().reshape(n, 32, cout // 32, h, w)
        rstd2 = (1.0 / (og.var(dim=(2, 3, 4), unbiased=False) + eps).sqrt()).repeat_interleave(cout // 32, dim=1).float()
        ok &= report("gn-fused conv: statistics of its own output (rstd)", s2, rstd2, rtol=5e-5, atol_scale=1e-6)
    return ok


def case_gn_fused_linear(n, P, C, cout, silu=False, eps=1e-6, force=0, seed=0):
    """SpatialTransformer norm -> proj_in (attention.py:399-404) fused: token-matrix producer (a Linear with residual, like
    proj_out) leaves statistics, the consumer Linear applies the GroupNorm in shared memory."""
    import torch
    from leftrefill_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(seed)
    M = n * P
    a = torch.randn(M, 64, generator=g).cuda().half()
    w0 = (torch.randn(C, 64, generator=g) / 8.0).cuda()
    r0 = torch.randn(M, C, generator=g).cuda().half()
    y, st = ops.gn_linear(a, ops.repack_linear(w0), P, bias=torch.randn(C, generator=g).cuda(), residual=r0,
                          want_stats=True)
    assert st is not None, "producer did not leave statistics"
    gamma = torch.randn(C, generator=g).cuda()
    beta = torch.randn(C, generator=g).cuda()
    scale, shift = ops.gn_finalize(st, gamma, beta, eps, n)
    yg = y.double().reshape(n, P, 32, C // 32)
    mean = yg.mean(dim=(1, 3))
    rstd = 1.0 / (yg.var(dim=(1, 3), unbiased=False) + eps).sqrt()
    sc_ref = (rstd.repeat_interleave(C // 32, dim=1) * gamma.double()[None]).float()
    ok = report(f"gn-finalize (token matrix) scale n={n} P={P} C={C}", scale, sc_ref, rtol=2e-5, atol_scale=1e-6)
    w1 = (torch.randn(cout, C, generator=g) / C ** 0.5).cuda()
    b1 = torch.randn(cout, generator=g).cuda()
    out = ops.gn_linear(y, ops.repack_linear(w1), P, gn=(scale, shift), silu=silu, bias=b1, force_block_n=force)
    torch.cuda.synchronize()
    xn = y.float().reshape(n, P, C) * scale[:, None, :] + shift[:, None, :]
    if silu:
        xn = torch.nn.functional.silu(xn)
    ref = xn.half().float().reshape(M, C) @ w1.half().float().t() + b1
    ok &= report(f"gn-fused linear n={n} P={P} C={C}->{cout} silu={silu} bn={force}", out, ref)
    return ok


def case_ln(M, C, seed=0):
    import torch
    import torch.nn.functional as F
    from leftrefill_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = (torch.randn(M, C, generator=g) * 2 + 0.3).cuda().half()
    gamma = torch.randn(C, generator=g).cuda()
    beta = torch.randn(C, generator=g).cuda()
    out = ops.layernorm(x, gamma, beta)
    torch.cuda.synchronize()
    ref = F.layer_norm(x.float(), (C,), gamma, beta, 1e-5)
    return report(f"layernorm M={M} C={C}", out, ref)


def case_time(kind):
    """Rough kernel timings at UNet sizes (CUDA events, L2-cold not enforced: bring-up only)."""
    import torch
    from leftrefill_b200 import ops
    torch.manual_seed(0)

    def timeit(fn, flops, name, iters=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"TIME [{name}] {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s", flush=True)

    if kind == "linear":
        for (M, K, Nn) in [(65536, 320, 960), (65536, 320, 320), (65536, 1280, 320), (16384, 640, 1920),
                           (4096, 1280, 3840), (1024, 1280, 1280), (65536, 128, 320)]:
            a = torch.randn(M, K, device="cuda").half()
            w = torch.randn(Nn, K, device="cuda").half()
            timeit(lambda: ops.linear(a, w), 2.0 * M * K * Nn, f"linear {M}x{K}x{Nn}")
            timeit(lambda: ops.linear(a, w, force_block_n=1000), 2.0 * M * K * Nn, f"linear {M}x{K}x{Nn} cg1")
        M, K, Nn = 65536, 320, 1280
        a = torch.randn(M, K, device="cuda").half()
        w = torch.randn(2 * Nn, K, device="cuda").half()
        timeit(lambda: ops.linear(a, w, geglu=True), 2.0 * M * K * 2 * Nn, f"geglu {M}x{K}x{2 * Nn}")
    elif kind == "conv":
        for (n, h, w_, c, co) in [(8, 64, 128, 320, 320), (8, 32, 64, 640, 640), (8, 16, 32, 1280, 1280),
                                  (8, 8, 16, 1280, 1280), (8, 64, 128, 640, 320), (8, 32, 64, 1280, 640)]:
            x = torch.randn(n, h, w_, c, device="cuda").half()
            wt = torch.randn(co, 9 * c, device="cuda").half() * 0.01
            b = torch.zeros(co, device="cuda")
            timeit(lambda: ops.conv3x3(x, wt, bias=b), 2.0 * n * h * w_ * 9 * c * co, f"conv {n}x{h}x{w_} {c}->{co}")
            for f in (1000, 2160, 2256):
                timeit(lambda: ops.conv3x3(x, wt, bias=b, force_block_n=f), 2.0 * n * h * w_ * 9 * c * co,
                       f"conv {n}x{h}x{w_} {c}->{co} force={f}")
    elif kind == "attn":
        for (b, hd, t, tk) in [(8, 5, 8192, 8192), (8, 10, 2048, 2048), (8, 20, 512, 512), (8, 5, 8192, 77),
                               (8, 20, 128, 128)]:
            q = torch.randn(b, t, hd * 64, device="cuda").half()
            k = torch.randn(b, tk, hd * 64, device="cuda").half()
            v = torch.randn(b, tk, hd * 64, device="cuda").half()
            timeit(lambda: ops.attention(q, k, v, hd), 4.0 * b * hd * t * tk * 64, f"attn b={b} h={hd} {t}x{tk}",
                   iters=5)
    return True


CASES = {
    "linear_min": lambda: case_linear(128, 64, 32, force=32),
    "linear_k256": lambda: case_linear(128, 256, 128, force=128),
    "linear_bn64": lambda: case_linear(512, 320, 320, force=64),
    "linear_bn160": lambda: case_linear(512, 320, 320, force=160),
    "linear_bn256": lambda: case_linear(384, 640, 512, force=256),
    "linear_2sm_256": lambda: case_linear(1024, 640, 512, bias=True, residual=True, force=2256),
    "linear_2sm_160": lambda: case_linear(1000, 320, 320, bias=True, residual=True, force=2160),
    "linear_2sm_odd": lambda: case_linear(384, 320, 96, bias=True, force=2096),
    "linear_2sm_geglu": lambda: case_linear(512, 320, 1280, bias=True, geglu=True, force=2256),
    "linear_2sm_big": lambda: case_linear(8192, 1280, 1280, bias=True, residual=True, force=2000),
    "linear_ragged": lambda: case_linear(1000, 320, 320, bias=True, residual=True),
    "linear_geglu": lambda: case_linear(512, 320, 1280, bias=True, geglu=True),
    "linear_n4": lambda: case_linear(300, 128, 4, bias=True),
    "linear_big": lambda: case_linear(8192, 1280, 1280, bias=True, residual=True),
    "conv_small": lambda: case_conv(2, 8, 16, 64, 64),
    "conv_mid": lambda: case_conv(2, 16, 32, 128, 192, bias_img=True, residual=True),
    "conv_concat": lambda: case_conv(2, 16, 32, 128, 128, c1=64),
    "conv_w128": lambda: case_conv(1, 8, 128, 64, 64),
    "conv_odd": lambda: case_conv(3, 24, 40, 64, 96),
    "conv_tiny": lambda: case_conv(4, 4, 8, 64, 64),
    "conv_2sm": lambda: case_conv(2, 16, 32, 128, 192, bias_img=True, residual=True, force=2000),
    "conv_2sm_concat": lambda: case_conv(2, 16, 32, 128, 128, c1=64, force=2128),
    "conv_2sm_odd": lambda: case_conv(3, 24, 40, 64, 96, force=2096),
    "conv_2sm_s2": lambda: case_conv(2, 64, 128, 64, 64, stride=2, force=2064),
    "conv_halo_ragged": lambda: case_conv(2, 17, 9, 64, 96),
    "conv_halo_big": lambda: case_conv(2, 64, 128, 320, 320, bias_img=True, residual=True),
    "conv_halo_concat": lambda: case_conv(1, 32, 64, 128, 160, c1=192, force=2160),
    "conv_halo_1cta": lambda: case_conv(3, 40, 24, 64, 64, residual=True, force=1064),
    "conv_splitk": lambda: case_conv(8, 8, 16, 128, 256, bias_img=True, residual=True),
    "conv_splitk_concat": lambda: case_conv(4, 8, 16, 64, 192, c1=128),
    "conv_splitk_s2": lambda: case_conv(2, 16, 32, 64, 128, stride=2, residual=True),
    "conv_s2": lambda: case_conv(2, 16, 32, 64, 64, stride=2),
    "conv_s2_big": lambda: case_conv(2, 64, 128, 64, 64, stride=2),
    "conv_cout4": lambda: case_conv(2, 16, 32, 64, 4),
    "upconv_small": lambda: case_upconv(2, 8, 16, 64, 64),
    "upconv_ragged": lambda: case_upconv(1, 9, 13, 64, 96),
    "upconv_big": lambda: case_upconv(2, 32, 64, 640, 640),
    "attn_min": lambda: case_attn(1, 1, 128, 128),
    "attn_q384": lambda: case_attn(2, 1, 384, 256),
    "attn_k4": lambda: case_attn(1, 2, 256, 512),
    "attn_tiles": lambda: case_attn(2, 2, 512, 512),
    "attn_fused": lambda: case_attn(2, 5, 256, 256, fused_qkv=True),
    "attn_cross": lambda: case_attn(2, 2, 256, 77),
    "attn_ragged": lambda: case_attn(1, 3, 200, 300),
    "attn_long": lambda: case_attn(1, 1, 1024, 4096),
    "attn_bigrange": lambda: case_attn_bigrange(1, 2, 256, 2048),
    "attn_bigrange_ragged": lambda: case_attn_bigrange(2, 1, 200, 1000, logit_std=24.0, ramp=80.0),
    "gn_bigmean": lambda: case_gn_bigmean(2, 16, 32, 320),
    "gn_bigmean_twopass": lambda: case_gn_bigmean(1, 64, 128, 320, ratio=60.0),
    "geglu_big": lambda: case_geglu_big(512, 320, 640),
    "gnf_conv": lambda: case_gn_fused_conv(2, 16, 32, 64, 96),
    "gnf_conv_big": lambda: case_gn_fused_conv(2, 64, 128, 320, 320, residual=True),
    "gnf_conv_concat": lambda: case_gn_fused_conv(2, 32, 64, 128, 160, c1=192, force=2160),
    "gnf_conv_1cta": lambda: case_gn_fused_conv(3, 40, 24, 64, 64, residual=True, force=1064),
    "gnf_conv_ragged": lambda: case_gn_fused_conv(2, 17, 33, 64, 96),
    "gnf_conv_nosilu": lambda: case_gn_fused_conv(1, 16, 8, 64, 64, silu=False, eps=1e-6),
    "gnf_conv_bigmean": lambda: case_gn_fused_conv(2, 32, 32, 320, 64, bigmean=True),
    "gnf_conv_cout4": lambda: case_gn_fused_conv(2, 16, 32, 320, 4),
    "gnf_linear": lambda: case_gn_fused_linear(2, 512, 320, 320),
    "gnf_linear_1cta": lambda: case_gn_fused_linear(3, 128, 64, 96, force=1096),
    "gnf_linear_2sm": lambda: case_gn_fused_linear(2, 2048, 640, 640, force=2000),
    "gnf_linear_silu": lambda: case_gn_fused_linear(1, 256, 128, 64, silu=True, eps=1e-5),
    "gn": lambda: case_gn(2, 16, 32, 320),
    "gn_concat": lambda: case_gn(2, 8, 16, 1280, c1=640),
    "gn_nosilu": lambda: case_gn(3, 8, 8, 64, silu=False, eps=1e-6),
    "ln": lambda: case_ln(1000, 320),
    "ln_1280": lambda: case_ln(256, 1280),
    "time_linear": lambda: case_time("linear"),
    "time_conv": lambda: case_time("conv"),
    "time_attn": lambda: case_time("attn"),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default=None)
    ap.add_argument("--only", default=None, help="comma separated prefixes")
    args = ap.parse_args()
    if args.case:
        ok = CASES[args.case]()
        sys.exit(0 if ok else 1)
    results = {}
    for name in CASES:
        if args.only and not any(name.startswith(p) for p in args.only.split(",")):
            continue
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", name], timeout=int(os.environ.get("LR_CASE_TIMEOUT", "120")),
                               capture_output=True, text=True)
            out = (r.stdout + r.stderr).strip().splitlines()
            tail = [l for l in out if l.startswith(("PASS", "FAIL", "TIME", "lr_b200"))] or out[-6:]
            print(f"== {name} rc={r.returncode} ({time.time() - t0:.1f}s)")
            for l in tail[-12:]:
                print("   " + l)
            results[name] = r.returncode
        except subprocess.TimeoutExpired:
            print(f"== {name} TIMEOUT")
            results[name] = -9
        sys.stdout.flush()
    bad = [k for k, v in results.items() if v != 0]
    print(f"SUMMARY: {len(results) - len(bad)}/{len(results)} ok; failing: {bad}")


if __name__ == "__main__":
    main()
