"""GPU parity tests of the native first-stage decoder (lr_vae_* through leftrefill_b200.AutoencoderKL.decode) against the
reference Decoder's golden outputs (oracle/make_golden.py --only-vae) and the oracle at BASELINE sizes.

Bars, as for the UNet: rel-RMS <= 4e-3 and max |err| <= 1.5e-2 * max|ref| against the fp32 reference, and not worse than
1.25x the error of the oracle run under torch.autocast (the reference decodes under autocast, test_inpainting.py:126)."""
import pytest
import torch

from helpers import err_stats, load_golden
from oracle import vae_oracle as V

pytestmark = pytest.mark.gpu

REL_RMS_BAR, MAX_ABS_BAR, FLOOR_FACTOR = 4e-3, 1.5e-2, 1.25


def _build(cfg, seed=0):
    import leftrefill_b200 as lr
    sd = V.make_state_dict(cfg, seed=seed)
    m = lr.AutoencoderKL(ddconfig={k: v for k, v in cfg.items() if k != "embed_dim"}, embed_dim=cfg["embed_dim"])
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval(), sd


def _floor(sd, cfg, z, scale):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sdc = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad(), torch.autocast("cuda"):
        return V.decode(sdc, cfg, z.cuda(), scale_factor=scale).float()


def _assert_parity(got, ref, floor=None):
    s = err_stats(got, ref)
    assert s["finite"], s
    assert s["rel_rms"] <= REL_RMS_BAR, s
    assert s["max_abs"] <= MAX_ABS_BAR * s["ref_max"], s
    if floor is not None:
        f = err_stats(floor, ref)
        assert s["rms"] <= FLOOR_FACTOR * f["rms"], (s, f)
    return s


def test_vae_small_vs_reference_golden():
    g = load_golden("vae_small.npz")
    m, sd = _build(V.SMALL_CFG)
    z = torch.tensor(g["z"])
    scale = float(g["scale_factor"])
    y = m.decode(z.cuda(), z_scale=1.0 / scale)
    assert y.dtype == torch.float32 and tuple(y.shape) == g["out"].shape
    _assert_parity(y, g["out"], _floor(sd, V.SMALL_CFG, z, scale))
    # batch independence: one latent alone gives the same image, bit for bit
    y0 = m.decode(z[:1].cuda(), z_scale=1.0 / scale)
    assert torch.equal(y0, y[:1])


def test_vae_full_config_vs_reference_golden():
    g = load_golden("vae_full_8x16.npz")
    m, sd = _build(V.DEFAULT_CFG)
    z = torch.tensor(g["z"])
    scale = float(g["scale_factor"])
    y = m.decode(z.cuda(), z_scale=1.0 / scale)
    _assert_parity(y, g["out"], _floor(sd, V.DEFAULT_CFG, z, scale))


def test_vae_full_size_vs_oracle():
    """BASELINE size: the SD2 decoder on 64x128 latents -> 512x1024 images (mid-block attention over 8192 tokens with
    d = 512), two latents, against the fp32 oracle on the GPU."""
    m, sd = _build(V.DEFAULT_CFG)
    g = torch.Generator().manual_seed(5)
    z = torch.randn(2, 4, 64, 128, generator=g) * V.SCALE_FACTOR * 4.0
    y = m.decode(z.cuda(), z_scale=1.0 / V.SCALE_FACTOR)
    assert tuple(y.shape) == (2, 3, 512, 1024) and torch.isfinite(y).all()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sdc = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        ref = V.decode(sdc, V.DEFAULT_CFG, z[:1].cuda(), scale_factor=V.SCALE_FACTOR)
        with torch.autocast("cuda"):
            floor = V.decode(sdc, V.DEFAULT_CFG, z[:1].cuda(), scale_factor=V.SCALE_FACTOR).float()
    _assert_parity(y[:1], ref, floor)
    del ref, floor
    torch.cuda.empty_cache()


def test_vae_weight_update_and_autocast():
    g = load_golden("vae_small.npz")
    m, _ = _build(V.SMALL_CFG)
    z = torch.tensor(g["z"]).cuda()
    y0 = m.decode(z)
    with torch.no_grad():
        m.decoder.conv_out.weight.mul_(2.0)
    y1 = m.decode(z)
    b = m.decoder.conv_out.bias.detach()[None, :, None, None]
    assert torch.allclose(y1 - b, 2 * (y0 - b), rtol=5e-3, atol=1e-2)
    with torch.autocast("cuda"):
        assert m.decode(z).dtype == torch.float16
    with pytest.raises(NotImplementedError):
        m.encode(torch.zeros(1, 3, 64, 64, device="cuda"))
