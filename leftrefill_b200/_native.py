"""ctypes binding of liblr_b200.so (the C ABI declared in include/lr_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, this raises.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int64, c_longlong, c_size_t, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
# LR_B200_LIB: A/B experiments against another build of the same ABI (tests/gpu_time_gemm.py); normal use never sets it
LIB_PATH = os.environ.get("LR_B200_LIB") or os.path.join(_HERE, "liblr_b200.so")


ABI_VERSION = 3  # LR_B200_ABI_VERSION of include/lr_b200.h


class LRError(RuntimeError):
    pass


class UNetCfg(Structure):
    _fields_ = [
        ("in_channels", c_int),
        ("model_channels", c_int),
        ("out_channels", c_int),
        ("num_levels", c_int),
        ("channel_mult", c_int * 8),
        ("num_res_blocks", c_int * 8),
        ("attention_ds", c_int * 8),
        ("n_attention_ds", c_int),
        ("num_head_channels", c_int),
        ("transformer_depth", c_int),
        ("context_dim", c_int),
        ("use_linear_in_transformer", c_int),
        ("view_num", c_int),
        ("concat_target", c_int),
        ("use_sep", c_int),
    ]


class VaeCfg(Structure):
    _fields_ = [
        ("ch", c_int),
        ("out_ch", c_int),
        ("num_levels", c_int),
        ("ch_mult", c_int * 8),
        ("num_res_blocks", c_int),
        ("z_channels", c_int),
        ("embed_dim", c_int),
    ]


# name -> (restype, argtypes); every symbol include/lr_b200.h declares
SIGNATURES = {
    "lr_abi_version": (c_int, []),
    "lr_last_error": (c_char_p, []),
    "lr_launch_count": (c_longlong, []),
    "lr_launch_count_reset": (None, []),
    "lr_launch_count_add": (None, [c_longlong]),
    "lr_debug_read_trace": (c_int, [c_void_p, c_longlong, c_int]),
    "lr_unet_create": (c_int, [POINTER(UNetCfg), POINTER(c_void_p)]),
    "lr_unet_destroy": (None, [c_void_p]),
    "lr_unet_num_weights": (c_int, [c_void_p]),
    "lr_unet_weight_name": (c_char_p, [c_void_p, c_int]),
    "lr_unet_weight_shape": (c_int, [c_void_p, c_int, POINTER(c_int64)]),
    "lr_unet_set_weight": (c_int, [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int, c_void_p]),
    "lr_unet_missing_weights": (c_int, [c_void_p]),
    "lr_unet_set_context": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "lr_unet_set_c_input": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "lr_unet_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                                c_void_p]),
    "lr_unet_forward_cfg_pair": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "lr_unet_set_profiling": (c_int, [c_void_p, c_int]),
    "lr_unet_read_profile": (c_int, [c_void_p, POINTER(c_double), POINTER(c_double), POINTER(c_int)]),
    "lr_unet_num_steps": (c_int, [c_void_p]),
    "lr_unet_step_info": (c_int, [c_void_p, c_int, POINTER(c_double), POINTER(c_double), POINTER(c_int), c_char_p,
                                  c_int]),
    "lr_unet_plan_generation": (c_longlong, [c_void_p]),
    "lr_vae_create": (c_int, [POINTER(VaeCfg), POINTER(c_void_p)]),
    "lr_vae_destroy": (None, [c_void_p]),
    "lr_vae_num_weights": (c_int, [c_void_p]),
    "lr_vae_weight_name": (c_char_p, [c_void_p, c_int]),
    "lr_vae_weight_shape": (c_int, [c_void_p, c_int, POINTER(c_int64)]),
    "lr_vae_set_weight": (c_int, [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int, c_void_p]),
    "lr_vae_missing_weights": (c_int, [c_void_p]),
    "lr_vae_decode": (c_int, [c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_int, c_void_p]),
    "lr_vae_last_flops": (c_double, [c_void_p]),
    "lr_vae_device_bytes": (c_longlong, [c_void_p]),
    "lr_vae_num_steps": (c_int, [c_void_p]),
    "lr_ddim_update_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int64, c_void_p,
                                   c_void_p, c_void_p]),
    "lr_unet_last_flops": (c_double, [c_void_p]),
    "lr_unet_device_bytes": (c_longlong, [c_void_p]),
    "lr_ddim_update": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float, c_float, c_float,
                               c_float, c_int64, c_void_p, c_void_p, c_void_p]),
    "lr_linear_f16": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int,
                              c_void_p, c_int, c_int, c_int, c_void_p]),
    "lr_conv3x3_f16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "lr_conv_stats_rows": (c_longlong, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "lr_gn_conv3x3_f16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                  c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_int), c_int,
                                  c_void_p]),
    "lr_gn_linear_f16": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                 c_void_p, c_void_p, c_void_p, POINTER(c_int), c_int, c_void_p]),
    "lr_gn_finalize": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_void_p]),
    "lr_upsample2x_conv3x3_f16": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                          c_void_p, c_void_p]),
    "lr_attention_f16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p,
                                 c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "lr_groupnorm_scratch_bytes": (c_size_t, [c_int, c_int, c_int]),
    "lr_groupnorm_f16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                                 c_int, c_void_p, c_void_p, c_void_p]),
    "lr_layernorm_f16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p]),
    "lr_nchw_f32_to_nhwc_f16": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "lr_nhwc_f16_to_nchw_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "lr_repack_conv3x3_weight": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "lr_repack_linear_weight": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
}

_lib = None


def lib():
    """Loads the shared library (once). Raises LRError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LRError(
                f"{LIB_PATH} not found: build it with `python -m leftrefill_b200.build` "
                "(there is no CPU or PyTorch fallback for the UNet hot path)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        if handle.lr_abi_version() != ABI_VERSION:
            raise LRError("liblr_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def check(status, what=""):
    if status != 0:
        msg = lib().lr_last_error()
        raise LRError(f"{what}: {msg.decode() if msg else 'unknown error'}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


def current_stream():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)
