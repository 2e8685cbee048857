"""leftrefill_b200 — B200-native (sm_100a) implementation of LeftRefill's DDIM/UNet hot path.

Public surface (mirrors the reference class surface):
    UNetModel, MultiViewUnetModel, NVSUnetModel, ResBlock, Upsample, Downsample, TimestepEmbedSequential   (openaimodel.py)
    CrossAttention, BasicTransformerBlock, SpatialTransformer, FeedForward, GEGLU            (attention.py)
    DDIMSampler                                                                              (ddim.py)
    AutoencoderKL (decode), Decoder              first-stage decoder on the same kernels (autoencoder.py:87-90,
                                                 model.py:547-653)
    PromptContextCache, install_context_cache    memoised prompt contexts for the learned-prompt text encoders
                                                 (Refill_modules.py:160-191)
    install()  -> makes `ldm.modules.diffusionmodules.openaimodel.UNetModel`, `ldm.modules.attention.CrossAttention`
                  and `ldm.models.diffusion.ddim.DDIMSampler` resolve to the classes above, so yaml `target:` strings
                  and `from ldm... import ...` lines of the reference drivers keep working unchanged.
"""
import importlib
import sys
import types

__version__ = "0.1.0"

_LAZY = {
    "UNetModel": "unet", "MultiViewUnetModel": "unet", "NVSUnetModel": "unet", "ResBlock": "unet", "Upsample": "unet", "Downsample": "unet",
    "TimestepEmbedSequential": "unet", "TimestepBlock": "unet", "GroupNorm32": "unet",
    "CrossAttention": "attention", "MemoryEfficientCrossAttention": "attention", "BasicTransformerBlock": "attention",
    "SpatialTransformer": "attention", "FeedForward": "attention", "GEGLU": "attention",
    "DDIMSampler": "ddim",
    "PromptContextCache": "context_cache", "install_context_cache": "context_cache",
    "AutoencoderKL": "vae", "Decoder": "vae",
}


def __getattr__(name):
    if name in _LAZY:
        return getattr(importlib.import_module(f"{__name__}.{_LAZY[name]}"), name)
    raise AttributeError(name)


_PATCHES = {
    "ldm.modules.diffusionmodules.openaimodel": ["UNetModel", "ResBlock", "Upsample", "Downsample",
                                                 "TimestepEmbedSequential", "TimestepBlock"],
    "ldm.modules.diffusionmodules.multiview_unet": ["MultiViewUnetModel"],
    "ldm.modules.attention": ["CrossAttention", "MemoryEfficientCrossAttention", "BasicTransformerBlock",
                              "SpatialTransformer", "FeedForward", "GEGLU"],
    "ldm.models.diffusion.ddim": ["DDIMSampler"],
}


def install(verbose=False):
    """Route the reference's module paths to the native implementations.

    If the reference `ldm` package is importable (a LeftRefill checkout on sys.path) its modules are imported and the
    hot-path classes inside them are replaced in place, so everything else of the reference keeps working. Otherwise
    stand-in modules carrying only the replaced classes are registered under the same dotted names.
    Call this before the reference drivers import `ldm...` names with `from ... import ...`.
    """
    done = {}
    for modname, names in _PATCHES.items():
        try:
            mod = importlib.import_module(modname)
            how = "patched"
        except Exception:  # reference tree (or one of its dependencies) not importable: register a stand-in
            parts = modname.split(".")
            for i in range(1, len(parts) + 1):
                pkg = ".".join(parts[:i])
                if pkg not in sys.modules:
                    m = types.ModuleType(pkg)
                    m.__path__ = []
                    sys.modules[pkg] = m
                    if i > 1:
                        setattr(sys.modules[".".join(parts[:i - 1])], parts[i - 1], m)
            mod = sys.modules[modname]
            how = "stand-in"
        for n in names:
            setattr(mod, n, __getattr__(n))
        done[modname] = how
        if verbose:
            print(f"[leftrefill_b200] {modname}: {how} {names}")
    return done
