"""GPU parity tests of the op-level C ABI (lr_linear_f16, lr_conv3x3_f16, lr_attention_f16, lr_groupnorm_f16,
lr_layernorm_f16, lr_ddim_update) against fp32 references on identical fp16-rounded inputs.
Tolerance: |err| <= 2e-4*max(1, max|ref|) + 1e-3*|ref| element-wise (see gpu_diag_ops.report)."""
import pytest
import torch

import helpers
from helpers import O
import gpu_diag_ops as D

pytestmark = pytest.mark.gpu

PARITY_CASES = [k for k in D.CASES if not k.startswith("time_")]


@pytest.mark.parametrize("name", PARITY_CASES)
def test_op_parity(name):
    assert D.CASES[name]()


def test_empty_inputs_are_noops():
    from leftrefill_b200 import ops
    a = torch.zeros(0, 64, dtype=torch.float16, device="cuda")
    w = torch.zeros(32, 64, dtype=torch.float16, device="cuda")
    assert ops.linear(a, w).shape == (0, 32)
    x = torch.zeros(0, 8, 8, 64, dtype=torch.float16, device="cuda")
    assert ops.conv3x3(x, torch.zeros(64, 576, dtype=torch.float16, device="cuda")).shape == (0, 8, 8, 64)


def test_bad_arguments_raise():
    from leftrefill_b200 import ops, _native as N
    a = torch.zeros(8, 60, dtype=torch.float16, device="cuda")  # K not a multiple of 8
    w = torch.zeros(32, 60, dtype=torch.float16, device="cuda")
    with pytest.raises(N.LRError):
        ops.linear(a, w)


@pytest.mark.parametrize("kind", ["bias_res", "geglu", "plain"])
def test_lean_and_general_epilogue_are_bit_identical(kind, monkeypatch):
    """Which epilogue path a GEMM takes depends on its geometry (for tiny token counts: on the batch size), so both must
    produce the same bits - otherwise results would depend on how a batch is sharded over GPUs."""
    from leftrefill_b200 import ops
    g = torch.Generator().manual_seed(5)
    M, K, Nn = 1000, 320, 320
    a = (torch.randn(M, K, generator=g) * 0.7).cuda().half()
    geglu = kind == "geglu"
    w32 = (torch.randn(2 * Nn if geglu else Nn, K, generator=g) / K ** 0.5).cuda()
    wp = ops.repack_linear(w32, geglu=geglu)
    b = torch.randn(2 * Nn if geglu else Nn, generator=g).cuda() if kind != "plain" else None
    if geglu:
        b = ops.geglu_interleave(b)
    r = torch.randn(M, Nn, generator=g).cuda().half() if kind == "bias_res" else None
    lean = ops.linear(a, wp, bias=b, residual=r, geglu=geglu)
    monkeypatch.setenv("LR_NO_LEAN_EPI", "1")
    general = ops.linear(a, wp, bias=b, residual=r, geglu=geglu)
    assert torch.equal(lean, general)


def test_geglu_row_order_of_the_repack_kernel():
    """lr_repack_linear_weight(geglu=1) writes groups of four rows (value_2k, value_2k+1, gate_2k, gate_2k+1): the order
    ops.geglu_interleave documents and the GEGLU epilogue assumes."""
    from leftrefill_b200 import ops
    w = torch.randn(2 * 96, 64, generator=torch.Generator().manual_seed(3)).cuda()
    got = ops.repack_linear(w, geglu=True)
    assert torch.equal(got, ops.geglu_interleave(w).half())
    n = 96
    assert torch.equal(got[0], w[0].half()) and torch.equal(got[1], w[1].half())
    assert torch.equal(got[2], w[n].half()) and torch.equal(got[3], w[n + 1].half())
    assert torch.equal(got[4], w[2].half()) and torch.equal(got[6], w[n + 2].half())


@pytest.mark.parametrize("shape", [(64, 128, 320, 0), (32, 64, 640, 320)])
def test_persistent_groupnorm_is_batch_invariant_and_deterministic(shape):
    """Images above 1.3 MB take gn_persistent_kernel (one launch, per-chunk fp64 partials, grid barrier). Its result for
    an image must not depend on the batch it is normalised in, on the number of CTAs, or on the run (fixed reduction
    order, no atomics): bit-identical outputs for n = 1, 3, 8 and across repeats; and it must leave its barrier counters
    reusable (second call on the same scratch inside ops.groupnorm's allocator)."""
    from leftrefill_b200 import ops
    h, w, c0, c1 = shape
    g = torch.Generator().manual_seed(11)
    x0 = (torch.randn(8, h, w, c0, generator=g) * 2.0 + 0.5).cuda().half()
    x1 = torch.randn(8, h, w, c1, generator=g).cuda().half() if c1 else None
    gamma = torch.randn(c0 + c1, generator=g).cuda()
    beta = torch.randn(c0 + c1, generator=g).cuda()
    full = ops.groupnorm(x0, gamma, beta, 1e-5, silu=True, x1=x1)
    again = ops.groupnorm(x0, gamma, beta, 1e-5, silu=True, x1=x1)
    assert torch.equal(full, again)
    for n in (1, 3):
        part = ops.groupnorm(x0[:n].contiguous(), gamma, beta, 1e-5, silu=True,
                             x1=x1[:n].contiguous() if x1 is not None else None)
        assert torch.equal(part, full[:n])
    last = ops.groupnorm(x0[7:].contiguous(), gamma, beta, 1e-5, silu=True, x1=x1[7:].contiguous() if x1 is not None else None)
    assert torch.equal(last, full[7:])
    xs = torch.cat([x0.float(), x1.float()], dim=3) if x1 is not None else x0.float()
    ref = torch.nn.functional.silu(torch.nn.functional.group_norm(xs.permute(0, 3, 1, 2), 32, gamma, beta, 1e-5)).permute(0, 2, 3, 1)
    assert (full.float() - ref).abs().max().item() <= 2e-4 * ref.abs().max().item() + 1e-3 * ref.abs().max().item()


@pytest.mark.parametrize("cfg_scale,sigma", [(2.5, 0.0), (2.5, 0.37), (1.0, 0.2)])
def test_ddim_update_matches_oracle(cfg_scale, sigma):
    from leftrefill_b200 import ops
    g = torch.Generator().manual_seed(5)
    x, eu, ec, nz = (torch.randn(3, 4, 16, 32, generator=g) for _ in range(4))
    a_t, a_prev = 0.4321, 0.5678
    ref_prev, ref_x0 = O.ddim_step(x, eu, ec if cfg_scale != 1.0 else None, nz, cfg_scale, a_t, a_prev, sigma)
    got_prev, got_x0 = ops.ddim_update(x.cuda(), eu.cuda(), ec.cuda() if cfg_scale != 1.0 else None, nz.cuda(),
                                       cfg_scale, a_t, a_prev, sigma, (1 - a_t) ** 0.5)
    assert torch.allclose(got_prev.cpu(), ref_prev, rtol=1e-5, atol=1e-5)
    assert torch.allclose(got_x0.cpu(), ref_x0, rtol=1e-5, atol=1e-5)


def test_cross_attention_module_matches_oracle():
    """CrossAttention.forward stand-alone (attention.py:165-196) vs the oracle's _attention on the same weights."""
    from leftrefill_b200 import CrossAttention
    torch.manual_seed(0)
    m = CrossAttention(query_dim=320, context_dim=1024, heads=5, dim_head=64).cuda()
    x = torch.randn(2, 200, 320, device="cuda")
    ctx = torch.randn(2, 77, 1024, device="cuda")
    sd = {"a." + k: v.detach().half().float().cpu() for k, v in m.state_dict().items()}
    with torch.no_grad():
        ref_cross = O._attention(sd, "a.", x.half().float().cpu(), ctx.half().float().cpu(), 5)
        got_cross = m(x, context=ctx)
    s = helpers.err_stats(got_cross, ref_cross)
    assert s["finite"] and s["rel_rms"] < 2e-3 and s["max_abs"] < 5e-3 * max(1.0, s["ref_max"]), s
    ms = CrossAttention(query_dim=320, heads=5, dim_head=64).cuda()
    sd = {"a." + k: v.detach().half().float().cpu() for k, v in ms.state_dict().items()}
    with torch.no_grad():
        xr = x.half().float().cpu()
        ref_self = O._attention(sd, "a.", xr, xr, 5)
        got_self = ms(x)
    s = helpers.err_stats(got_self, ref_self)
    assert s["finite"] and s["rel_rms"] < 2e-3, s
