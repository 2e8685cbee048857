"""CPU: the oracle (oracle/unet_oracle.py) against the golden fixtures the UNMODIFIED reference produced
(oracle/make_golden.py, run in the build container). This is what pins the oracle."""
import numpy as np
import pytest
import torch

from helpers import O, load_golden

torch.set_grad_enabled(False)


@pytest.fixture(scope="module")
def small_sd():
    return O.make_state_dict(O.SMALL_CFG, seed=0)


def _close(got, ref, tol=5e-5):
    d = np.abs(np.asarray(got) - ref).max()
    assert d <= tol * max(1.0, np.abs(ref).max()), d


def test_param_walk_matches_reference_counts():
    spec = O.unet_spec(O.DEFAULT_CFG)
    assert len(spec) == 686                                     # 686 state-dict tensors
    assert sum(int(np.prod(s)) for _, s in spec) == 865_925_124  # 865.93 M parameters
    assert len({n for n, _ in spec}) == 686


def test_unet_small_vs_reference_golden(small_sd):
    g = load_golden("unet_small.npz")
    taps = {}
    y = O.unet_forward(small_sd, O.SMALL_CFG, torch.tensor(g["x"]), torch.tensor(g["t"]), torch.tensor(g["context"]),
                       taps=taps)
    _close(y.numpy(), g["out"])
    for k in [k for k in g if k.startswith("tap.")]:
        _close(taps[k[4:]].numpy(), g[k])


def test_unet_ragged_shape_vs_reference_golden(small_sd):
    g = load_golden("unet_small_ragged.npz")  # 24x40 latent, 50 context tokens
    y = O.unet_forward(small_sd, O.SMALL_CFG, torch.tensor(g["x"]), torch.tensor(g["t"]), torch.tensor(g["context"]))
    _close(y.numpy(), g["out"])


@pytest.mark.parametrize("name", ["multiview_v2.npz", "multiview_v3ct.npz"])
def test_multiview_vs_reference_golden(name):
    g = load_golden(name)
    sd = O.make_state_dict(O.SMALL_CFG, seed=1)
    y = O.unet_forward(sd, O.SMALL_CFG, torch.tensor(g["x"]), torch.tensor(g["t"]), torch.tensor(g["context"]),
                       view_num=int(g["view_num"]), concat_target=bool(g["concat_target"]))
    _close(y.numpy(), g["out"])


@pytest.mark.parametrize("tag,eta", [("eta0", 0.0), ("eta1", 1.0)])
def test_ddim_sampler_vs_reference_golden(small_sd, tag, eta):
    g = load_golden("ddim_small.npz")
    noises = [torch.tensor(n) for n in g[f"{tag}.noises"]]
    s = O.ddim_sample(small_sd, O.SMALL_CFG, torch.tensor(g["x_T"]), torch.tensor(g["c_concat"]),
                      torch.tensor(g["context"]), torch.tensor(g["uc_context"]), 4, eta, 2.5, noises)
    _close(s.numpy(), g[f"{tag}.samples"], tol=2e-4)


def test_schedule_vs_reference_golden():
    g = load_golden("ddim_small.npz")
    steps, a, ap, sg = O.make_schedule(50, 1.0, O.make_alphas_cumprod())
    assert (steps == g["sched50.timesteps"]).all() and steps[0] == 1 and steps[-1] == 981
    assert np.abs(a - g["sched50.alphas"]).max() < 1e-6
    assert np.abs(ap - g["sched50.alphas_prev"]).max() < 1e-6
    assert np.abs(sg - g["sched50.sigmas"]).max() < 1e-6


def test_timestep_embedding_edge_cases():
    e = O.timestep_embedding(torch.tensor([0, 1, 981]), 320)
    assert e.shape == (3, 320)
    assert torch.allclose(e[0, :160], torch.ones(160)) and torch.allclose(e[0, 160:], torch.zeros(160))
    assert O.timestep_embedding(torch.tensor([5]), 321).shape == (1, 321)  # odd dim gets a zero column


def test_ddim_multi_sampling_vs_reference_golden(small_sd):
    """List conditioning -> ddim_multi_sampling (ddim.py:104,146-222): two views, 4 steps, eta 1, cfg 2.5, the reference's
    recorded noises and `random` seed (oracle/make_golden.py --only-multi)."""
    import random
    g = load_golden("ddim_multi_small.npz")
    V, S = int(g["V"]), int(g["S"])
    noises = [torch.tensor(n) for n in g["noises"]]
    s = O.ddim_multi_sample(small_sd, O.SMALL_CFG, [torch.tensor(g[f"x_T{v}"]) for v in range(V)],
                            [torch.tensor(g[f"c_concat{v}"]) for v in range(V)],
                            [torch.tensor(g[f"context{v}"]) for v in range(V)], torch.tensor(g["uc_context"]), S, 1.0,
                            2.5, noises, random.Random(int(g["random_seed"])))
    _close(s.numpy(), g["samples"], tol=2e-4)


def test_nvs_use_sep_vs_reference_golden():
    """NVSUnetModel(use_sep=True) + c_input (inpainting_ldm/NVS_ldm.py:22-104): the oracle against the output of the
    unmodified reference class (oracle/make_golden.py --only-nvs; full config because the reference hard-codes the
    separator channel list for model_channels = 320), 16x32 latent -> feature widths 33 / 17 / 9 / 5."""
    g = load_golden("nvs_full_16x32.npz")
    cfg = O.DEFAULT_CFG
    assert O.sep_channels(cfg) == [9, 320, 640, 1280, 2560, 1920, 960]      # NVS_ldm.py:27
    sd = O.make_state_dict(cfg, seed=int(g["seed"]))
    sd.update(O.make_sep_tokens(cfg, seed=int(g["seed"])))
    x, t, ctx = torch.tensor(g["x"]), torch.tensor(g["t"]), torch.tensor(g["context"])
    y = O.nvs_unet_forward(sd, cfg, x, t, ctx, use_sep=True, c_input=torch.tensor(g["c_input_half"]))
    _close(y.numpy(), g["out.sep_cin_half"])
    y = O.nvs_unet_forward(sd, cfg, x[:1], t[:1], ctx[:1], use_sep=False,
                           c_input=torch.tensor(g["c_input_half_nosep"])[:1])
    _close(y.numpy(), g["out.nosep_cin_half"][:1])


def test_vae_decoder_vs_reference_golden():
    """First-stage decode (autoencoder.py:87-90, model.py:547-653): the oracle against the unmodified reference Decoder's
    output (oracle/make_golden.py --only-vae), incl. the state-dict walk of the full SD2 VAE."""
    from oracle import vae_oracle as V
    g = load_golden("vae_small.npz")
    sd = V.make_state_dict(V.SMALL_CFG, seed=0)
    taps = {}
    y = V.decode(sd, V.SMALL_CFG, torch.tensor(g["z"])[:1], scale_factor=float(g["scale_factor"]), taps=taps)
    _close(y.numpy(), g["out"][:1])
    _close(taps["mid"].numpy(), g["mid"][:1])
    spec = V.decoder_spec(V.DEFAULT_CFG)
    assert len(spec) == 140 and sum(int(np.prod(s)) for _, s in spec) == 49_490_199
    g = load_golden("vae_full_8x16.npz")
    y = V.decode(V.make_state_dict(V.DEFAULT_CFG, seed=0), V.DEFAULT_CFG, torch.tensor(g["z"]),
                 scale_factor=float(g["scale_factor"]))
    _close(y.numpy(), g["out"])
