"""CPU: host-side logic and the C-ABI contract (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from helpers import O, ROOT, FakeLDM, load_golden

import leftrefill_b200
from leftrefill_b200 import _native as N


def test_library_loads_and_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "lr_b200.h")).read()
    declared = set(re.findall(r"\b(lr_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    lib = ctypes.CDLL(N.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in lr_b200.h but not exported"
    assert declared == set(N.SIGNATURES), "ctypes signature table and header disagree"
    assert N.lib().lr_abi_version() == N.ABI_VERSION == 3


def _engine_table(model):
    L, h = N.lib(), model.engine()
    out = {}
    for i in range(L.lr_unet_num_weights(h)):
        shp = (ctypes.c_int64 * 4)()
        nd = L.lr_unet_weight_shape(h, i, shp)
        out[L.lr_unet_weight_name(h, i).decode()] = tuple(shp[:nd])
    return out


@pytest.mark.parametrize("cfg", [O.SMALL_CFG, dict(O.SMALL_CFG, use_linear_in_transformer=False, transformer_depth=2,
                                                   num_res_blocks=[1, 2, 1, 1])])
def test_state_dict_and_engine_weight_table_match_the_reference_layout(cfg):
    m = leftrefill_b200.UNetModel(**cfg)
    spec = O.unet_spec(cfg)
    sd = m.state_dict()
    assert list(sd.keys()) == [n for n, _ in spec]              # same keys, same order as the reference module
    assert all(tuple(sd[n].shape) == tuple(s) for n, s in spec)
    table = _engine_table(m)
    assert table == {n: tuple(s) for n, s in spec}              # native weight table == reference state dict
    assert N.lib().lr_unet_missing_weights(m.engine()) == len(spec)


def test_unsupported_options_fail_loudly():
    with pytest.raises(NotImplementedError):
        leftrefill_b200.UNetModel(**dict(O.SMALL_CFG, num_head_channels=32))
    with pytest.raises(NotImplementedError):
        leftrefill_b200.UNetModel(**dict(O.SMALL_CFG, use_scale_shift_norm=True))
    with pytest.raises(NotImplementedError):
        leftrefill_b200.UNetModel(**dict(O.SMALL_CFG, use_spatial_transformer=False, context_dim=None))


def test_nvs_unet_and_multi_sampling_surface():
    """N4 rows of SURVEY §8f: NVSUnetModel (inpainting_ldm/NVS_ldm.py:22-104) is the plain UNet when use_sep is False;
    use_sep=True adds the `sep_token.<channels>` parameters (the reference's hard-coded list for model_channels 320) to
    the state dict AND to the engine's weight table; unknown semantics-changing kwargs raise; DDIMSampler exposes
    ddim_multi_sampling with the reference's keyword surface (ddim.py:146-158)."""
    import inspect
    m = leftrefill_b200.NVSUnetModel(**dict(O.SMALL_CFG, use_sep=False))
    assert list(m.state_dict().keys()) == [n for n, _ in O.unet_spec(O.SMALL_CFG)]
    assert m.sep_token is None
    ms = leftrefill_b200.NVSUnetModel(**dict(O.DEFAULT_CFG, use_sep=True))
    want = [9, 320, 640, 1280, 2560, 1920, 960]                      # NVS_ldm.py:27
    assert O.sep_channels(O.DEFAULT_CFG) == want
    assert sorted(ms.sep_token.keys()) == sorted(str(c) for c in want)
    assert all(ms.sep_token[str(c)].shape == (c,) for c in want)
    table = _engine_table(ms)
    assert [n for n in ms.engine_weight_names() if n.startswith("sep_token.")] == [f"sep_token.{c}" for c in want]
    assert set(table) == set(ms.state_dict().keys())
    small = leftrefill_b200.NVSUnetModel(**dict(O.SMALL_CFG, use_sep=True))
    assert sorted(int(k) for k in small.sep_token.keys()) == sorted(O.sep_channels(O.SMALL_CFG))
    with pytest.raises(NotImplementedError):
        leftrefill_b200.MultiViewUnetModel(**dict(O.SMALL_CFG), view_num=2, no_rearrange_selfattn=True)
    with pytest.raises(NotImplementedError):
        leftrefill_b200.MultiViewUnetModel(**dict(O.SMALL_CFG), view_num=2, use_sep=True)
    sig = inspect.signature(leftrefill_b200.DDIMSampler.ddim_multi_sampling)
    for kw in ("cond", "shape", "x_T", "timesteps", "temperature", "noise_dropout", "unconditional_guidance_scale",
               "unconditional_conditioning", "ucg_schedule", "mask", "x0", "callback", "img_callback"):
        assert kw in sig.parameters, kw


def test_model_copy_and_pickle_drop_the_engine_handle():
    import copy
    import pickle
    m = leftrefill_b200.UNetModel(**O.SMALL_CFG)
    m.engine()                                   # a live ctypes handle
    m2 = copy.deepcopy(m)
    assert m2._engine is None and m._engine is not None
    m3 = pickle.loads(pickle.dumps(m))
    assert m3._engine is None
    assert list(m3.state_dict().keys()) == list(m.state_dict().keys())


class _CountingEmbedder(torch.nn.Module):
    """Stands in for PromptCLIPEmbedder (Refill_modules.py:100-191): text -> [B, 77, C], counts the strings it encodes."""

    def __init__(self, deep=0):
        super().__init__()
        self.special_embeddings = torch.nn.Embedding(4, 8)
        self.encoded = []
        self.deep = deep

    def forward(self, text):
        def one(s):
            g = torch.Generator().manual_seed(sum(map(ord, s)) + 1)
            return torch.randn(77, 8, generator=g) + self.special_embeddings.weight.sum()
        if self.deep:
            self.encoded += [tuple(layer[i] for layer in text) for i in range(len(text[0]))]
            return torch.stack([torch.stack([one(layer[i]) for layer in text]) for i in range(len(text[0]))])
        self.encoded += list(text)
        return torch.stack([one(s) for s in text])

    def encode(self, text):
        return self(text)


def test_prompt_context_cache():
    """contexts are memoised per unique prompt string; the encoder sees each string once; parameter
    updates and invalidate() drop the cache; the deep-prompt layout [n_layer][B] is keyed per sample."""
    from leftrefill_b200 import PromptContextCache, install_context_cache
    emb = _CountingEmbedder()
    ref = _CountingEmbedder()
    ref.load_state_dict(emb.state_dict())
    for q in list(emb.parameters()) + list(ref.parameters()):
        q.requires_grad_(False)                                  # frozen encoder, as at inference (freeze(), :152-155)
    c = PromptContextCache(emb)
    p = "<left> <right> a photo"
    a = c([p] * 4)
    assert a.shape == (4, 77, 8) and emb.encoded == [p]
    assert torch.equal(a, ref([p] * 4))
    u = c.encode([""] * 4)                                       # get_unconditional_conditioning, ref_inpainting_ldm.py:35
    assert emb.encoded == [p, ""] and torch.equal(u, ref([""] * 4))
    b = c([p, "", "other", p])
    assert emb.encoded == [p, "", "other"] and torch.equal(b, ref([p, "", "other", p]))
    assert c.hits == 3 + 3 + 3 and c.misses == 3 and c.encoder_calls == 3
    with torch.no_grad():
        emb.special_embeddings.weight.add_(1.0)                  # optimizer step: version counter moves
        ref.special_embeddings.weight.add_(1.0)
    assert torch.equal(c([p]), ref([p])) and emb.encoded[-1] == p and len(emb.encoded) == 4
    emb.special_embeddings.weight.data.mul_(2.0)                 # invisible to the version counter
    ref.special_embeddings.weight.data.mul_(2.0)
    c.invalidate()
    assert torch.equal(c([p]), ref([p])) and len(emb.encoded) == 5
    # deep prompt: [n_layer][B] lists, output [B, n_layer, L, C] (Refill_modules.py:162-171)
    d = _CountingEmbedder(deep=3)
    with torch.enable_grad():                                    # trainable embeddings + grad mode: bypass the cache
        n0 = len(d.encoded)
        PromptContextCache(d)([["a"], ["a1"], ["a2"]])
        PromptContextCache(d)([["a"], ["a1"], ["a2"]])
        assert len(d.encoded) == n0 + 2
    d = _CountingEmbedder(deep=3).requires_grad_(False)
    cd = PromptContextCache(d)
    text = [["a", "b", "a"], ["a1", "b1", "a1"], ["a2", "b2", "a2"]]
    z = cd(text)
    assert z.shape == (3, 3, 77, 8) and d.encoded == [("a", "a1", "a2"), ("b", "b1", "b2")]
    assert torch.equal(z[0], z[2]) and not torch.equal(z[0], z[1])

    class LDM:
        cond_stage_model = emb
    ldm = LDM()
    w = install_context_cache(ldm)
    assert ldm.cond_stage_model is w and install_context_cache(ldm) is w
    assert w.special_embeddings is emb.special_embeddings        # attributes stay reachable through the wrapper


def test_vae_decoder_weight_table_matches_the_reference_layout():
    """N2: AutoencoderKL (decode half) owns the reference's parameter names / shapes (autoencoder.py:33-35,
    model.py:547-606) and the native decoder's weight table is the same list, in the same order."""
    from oracle import vae_oracle as V
    for cfg in (V.SMALL_CFG, V.DEFAULT_CFG):
        m = leftrefill_b200.AutoencoderKL(ddconfig={k: v for k, v in cfg.items() if k != "embed_dim"},
                                          embed_dim=cfg["embed_dim"])
        spec = V.decoder_spec(cfg)
        assert list(m.state_dict().keys()) == [n for n, _ in spec]
        assert all(tuple(m.state_dict()[n].shape) == tuple(s) for n, s in spec)
        assert m.engine_weight_names() == [n for n, _ in spec]
        assert N.lib().lr_vae_missing_weights(m.engine()) == len(spec)
    with pytest.raises(N.LRError):
        m.decode(torch.zeros(1, 4, 8, 8))
    with pytest.raises(NotImplementedError):
        leftrefill_b200.Decoder(ch=32, out_ch=3, ch_mult=(1, 2), num_res_blocks=1, attn_resolutions=[16], in_channels=3,
                                resolution=64, z_channels=4)


def test_no_cpu_fallback():
    m = leftrefill_b200.UNetModel(**O.SMALL_CFG)
    with pytest.raises(N.LRError):
        m(torch.zeros(1, 9, 16, 32), torch.zeros(1, dtype=torch.long), context=torch.zeros(1, 77, 256))


def test_error_reporting_through_the_abi():
    L = N.lib()
    h = ctypes.c_void_p()
    cfg = N.UNetCfg()
    cfg.model_channels, cfg.num_levels, cfg.num_head_channels, cfg.transformer_depth = 60, 1, 64, 1
    assert L.lr_unet_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert b"model_channels" in L.lr_last_error()
    with pytest.raises(N.LRError):
        N.check(L.lr_unet_create(None, ctypes.byref(h)), "create")


def test_install_routes_reference_module_paths():
    leftrefill_b200.install()
    import importlib
    om = importlib.import_module("ldm.modules.diffusionmodules.openaimodel")
    at = importlib.import_module("ldm.modules.attention")
    dd = importlib.import_module("ldm.models.diffusion.ddim")
    assert om.UNetModel is leftrefill_b200.UNetModel
    assert at.CrossAttention is leftrefill_b200.CrossAttention
    assert dd.DDIMSampler is leftrefill_b200.DDIMSampler
    # the plugin mechanism of the reference: instantiate_from_config resolves `target` with importlib (ldm/util.py:71-86)
    target = "ldm.modules.diffusionmodules.openaimodel.UNetModel"
    module, cls = target.rsplit(".", 1)
    assert getattr(importlib.import_module(module), cls) is leftrefill_b200.UNetModel


def test_ddim_schedule_matches_reference_golden():
    g = load_golden("ddim_small.npz")
    s = leftrefill_b200.DDIMSampler(FakeLDM(None, torch.device("cpu")))
    s.make_schedule(50, ddim_eta=1.0, verbose=False)
    assert (s.ddim_timesteps == g["sched50.timesteps"]).all()
    assert np.abs(s.ddim_alphas - g["sched50.alphas"]).max() < 1e-6
    assert np.abs(s.ddim_alphas_prev - g["sched50.alphas_prev"]).max() < 1e-6
    assert np.abs(s.ddim_sigmas - g["sched50.sigmas"]).max() < 1e-6
    assert np.abs(s.ddim_sqrt_one_minus_alphas - g["sched50.sqrt_one_minus_alphas"]).max() < 1e-6


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "leftrefill_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"


def test_geglu_interleave_is_the_documented_row_permutation():
    """ops.geglu_interleave (the host-side mirror of lr_repack_linear_weight(geglu=1), checked against the kernel in
    tests/test_ops_gpu.py): source rows [0, n) values, [n, 2n) gates -> groups (v_2k, v_2k+1, g_2k, g_2k+1)."""
    import torch
    from leftrefill_b200 import ops
    n = 10
    src = torch.arange(2 * n, dtype=torch.float32)
    out = ops.geglu_interleave(src)
    assert sorted(out.tolist()) == src.tolist()                       # a permutation
    for k in range(n // 2):
        assert out[4 * k:4 * k + 4].tolist() == [2 * k, 2 * k + 1, n + 2 * k, n + 2 * k + 1]
    w = torch.randn(2 * n, 3)
    assert torch.equal(ops.geglu_interleave(w)[:, 1], ops.geglu_interleave(w[:, 1].contiguous()))
