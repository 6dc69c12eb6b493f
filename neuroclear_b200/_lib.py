"""ctypes binding of libneuroclear_b200.so (the C ABI declared in include/neuroclear_b200.h).

The library is built in-tree by ``neuroclear_b200.build`` (nvcc, sm_100a).  There is no fallback: if the shared
object is missing, or a call fails, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libneuroclear_b200.so")

i32, i64, f32, f64, vp = C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_void_p
I3 = C.c_int32 * 3
U4 = C.c_uint64 * 4

# name -> (restype, argtypes); mirrors include/neuroclear_b200.h one to one
SIGNATURES = {
    "nc_abi_version": (C.c_int, []),
    "nc_last_error": (C.c_char_p, []),
    "nc_build_source_hash": (C.c_char_p, []),
    "nc_device_sm_count": (C.c_int, []),
    "nc_memcpy2d_h2d_async": (C.c_int, [vp, vp, i64, i64, i64, vp]),
    "nc_debug_set_max_ctas": (None, [i32]),
    "nc_debug_set_remainder_pairs": (None, [i32]),
    "nc_debug_set_disc_cluster": (None, [i32]),
    "nc_dice_geometry": (i64, [I3, i32, i32, I3, I3]),
    "nc_dice_extract_u16": (C.c_int, [vp, i32, i32, I3, I3, I3, i32, i32, i32, i64, i32, vp, vp]),
    "nc_dice_extract_u8": (C.c_int, [vp, i32, i32, I3, I3, I3, i32, i32, i32, i64, i32, vp, vp]),
    "nc_conv3d_k3_stats_rows": (i64, [i32, i32, i32, i32, i32, i32]),
    "nc_pack_weights_conv3d_cin1_k3": (C.c_int, [vp, vp, vp]),
    "nc_conv3d_cin1_k3_fwd": (C.c_int, [vp, vp, i32, i32, i32, i32, i32, vp, vp, vp]),
    "nc_packed_weight_bytes": (i64, [i32, i32, i32]),
    "nc_pack_weights_conv3d_k3": (C.c_int, [vp, i32, i32, vp, vp]),
    "nc_pack_weights_convT3d_k2s2": (C.c_int, [vp, i32, i32, vp, vp]),
    "nc_conv3d_k3_fwd": (C.c_int, [vp, vp, i32, i32, i32, i32, i32, vp, i32, vp, vp, vp]),
    "nc_pack_weights_conv3d_k3_dgrad": (C.c_int, [vp, i32, i32, vp, vp]),
    "nc_conv3d_k3_dgrad": (C.c_int, [vp, i32, i32, i32, i32, i32, vp, i32, vp, vp]),
    "nc_conv3d_wgrad_scratch_bytes": (i64, [i32, i32, i32, i32, i32, i32, i32]),
    "nc_conv3d_wgrad": (C.c_int, [vp, i32, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp]),
    "nc_space_to_depth_bf16": (C.c_int, [vp, i32, i32, i32, i32, i32, i32, i32, vp, vp]),
    "nc_pack_weights_convT3d_k2s2_dgrad": (C.c_int, [vp, i32, i32, vp, vp]),
    "nc_conv3d_k1_bf16": (C.c_int, [vp, i32, i32, i32, i32, i32, vp, i32, vp, vp]),
    "nc_colsum_bf16": (C.c_int, [vp, i32, i32, i32, i64, i32, vp, vp, vp]),
    "nc_cast_f16_bf16": (C.c_int, [vp, i32, i32, i64, i32, vp, i32, i32, vp]),
    "nc_bwd_scratch_bytes": (i64, [i32]),
    "nc_in_relu_apply_bf16": (C.c_int, [vp, vp, i32, i32, i32, i32, i32, vp, i32, i32, vp, vp]),
    "nc_in_relu_bwd": (C.c_int, [vp, vp, i32, i32, i32, i32, i32, i32, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp]),
    "nc_head_1x1_sigmoid_bwd": (C.c_int, [vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp]),
    "nc_conv3d_cin1_k3_wgrad": (C.c_int, [vp, vp, i32, i32, i32, i32, i32, vp, vp, vp]),
    "nc_im2col49": (C.c_int, [vp, i32, i32, i32, i32, i32, vp, vp]),
    "nc_col2im49": (C.c_int, [vp, i32, i32, i32, i32, vp, vp]),
    "nc_pack_weights_64": (C.c_int, [vp, i32, i32, vp, vp]),
    "nc_conv3d_tc_64": (C.c_int, [vp, i32, i32, i32, i32, i32, vp, i32, i32, vp, vp]),
    "nc_stencil64to1_fwd": (C.c_int, [vp, vp, i32, i32, i32, i32, vp, vp]),
    "nc_stencil64to1_bwd_data": (C.c_int, [vp, vp, i32, i32, i32, i32, vp, vp]),
    "nc_augment_crop_u16": (C.c_int, [vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, i32, vp, vp]),
    "nc_adam_step_multi": (C.c_int, [vp, i32, f32, f32, f32, f32, i32, vp]),
    "nc_patchgan_ws_floats": (i64, [i32, i32, i32, i32, i32]),
    "nc_patchgan_fwd": (C.c_int, [vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp]),
    "nc_patchgan_bwd": (C.c_int, [vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]),
    "nc_convT3d_k2s2_fwd": (C.c_int, [vp, vp, i32, i32, i32, i32, i32, vp, vp, i32, vp, i32, i32, vp]),
    "nc_in_stats_scratch_bytes": (i64, [i32, i32]),
    "nc_in_stats_finalize": (C.c_int, [vp, i32, i64, i32, i64, f32, vp, vp, vp]),
    "nc_in_relu_apply": (C.c_int, [vp, vp, i32, i32, i32, i32, i32, vp, i32, i32, vp, vp]),
    "nc_head_1x1_sigmoid_fwd": (C.c_int, [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp]),
    "nc_blend_gather_f32": (C.c_int, [vp, vp, vp, I3, I3, i32, i32, i32, i32, vp, vp]),
    "nc_select_init": (C.c_int, [U4, vp, vp]),
    "nc_select_histogram": (C.c_int, [vp, i64, i32, vp, vp, vp]),
    "nc_select_update": (C.c_int, [i32, vp, vp, vp]),
    "nc_percentile_lerp": (C.c_int, [vp, f64, f64, vp, vp, vp]),
    "nc_rescale_u16_crop": (C.c_int, [vp, i32, I3, I3, vp, i32, i32, vp, vp]),
    "nc_rescale_u8_crop": (C.c_int, [vp, i32, I3, I3, vp, i32, i32, vp, vp]),
    "nc_mip_fwd": (C.c_int, [vp, i32, i32, i32, i32, i32, i32, vp, vp, vp]),
    "nc_mip_bwd": (C.c_int, [vp, vp, i32, i32, i32, i32, vp, vp]),
    "nc_conv2d_k4_fwd": (C.c_int, [vp, vp, vp, i32, i32, i32, i32, i32, i32, f32, vp, vp]),
    "nc_conv2d_k4_dgrad": (C.c_int, [vp, vp, i32, i32, i32, i32, i32, i32, vp, vp]),
    "nc_conv2d_k4_wgrad": (C.c_int, [vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp]),
    "nc_in2d_lrelu_fwd": (C.c_int, [vp, i32, i32, f32, f32, vp, vp, vp]),
    "nc_in2d_lrelu_bwd": (C.c_int, [vp, vp, vp, i32, i32, f32, vp, vp]),
    "nc_lrelu_bwd": (C.c_int, [vp, vp, i64, f32, vp, vp]),
    "nc_loss_fwd": (C.c_int, [vp, vp, f32, i64, i32, vp, vp]),
    "nc_loss_bwd": (C.c_int, [vp, vp, f32, i64, i32, vp, vp, vp]),
    "nc_adam_step": (C.c_int, [vp, vp, vp, vp, i64, f32, f32, f32, f32, i32, vp]),
    "nc_unet_deconv_workspace_bytes": (i64, [i32, i32, i32, i32]),
    "nc_unet_deconv_workspace_init": (C.c_int, [vp, i64, i32, i32, i32, i32, vp]),
    "nc_unet_deconv_infer_cube": (C.c_int, [vp, i32, i32, i32, i32, vp, vp, i64, i32, vp, vp]),
    "nc_hist_match_scratch_bytes": (i64, [i64]),
    "nc_hist_match_f32": (C.c_int, [vp, vp, i64, vp, i64, vp, vp]),
    "nc_blend_gather_f64": (C.c_int, [vp, vp, vp, I3, I3, i32, i32, i32, i32, vp, vp]),
    "nc_amax_axis": (C.c_int, [vp, i32, i32, i32, i32, i32, i32, i32, vp, vp]),
    "nc_volume_moments": (C.c_int, [vp, i32, i64, vp, vp]),
    "nc_pairwise_sqdev_scratch_doubles": (i64, [i64]),
    "nc_pairwise_sqdev_sum": (C.c_int, [vp, i32, i64, f64, vp, vp, vp]),
    "nc_standardize_normalize_u8": (C.c_int, [vp, i32, i64, vp, vp, vp]),
    "nc_sqdiff_u8": (C.c_int, [vp, vp, i64, vp, vp]),
}

_lib = None


class NeuroclearError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises NeuroclearError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NeuroclearError(
                f"{LIB_PATH} is missing: build it with `python -m neuroclear_b200.build` "
                "(there is no CPU or PyTorch fallback for the CUDA hot path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.nc_abi_version() != 1:
            raise NeuroclearError("libneuroclear_b200.so ABI version mismatch; rebuild")
        if os.environ.get("NEUROCLEAR_REMAINDER_PAIRS", "1") == "0":      # A/B and debugging hook
            lib.nc_debug_set_remainder_pairs(0)
        _lib = lib
    return _lib


def check(rc: int):
    if rc != 0:
        raise NeuroclearError(load().nc_last_error().decode())


#: number of kernel launches issued through `call` (every int-returning entry point launches exactly one kernel)
LAUNCHES = 0


def call(name: str, *args):
    """Call an int-returning entry point (one kernel launch) and raise on error."""
    global LAUNCHES
    check(getattr(load(), name)(*args))
    LAUNCHES += 1


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device (or host) address of a torch tensor / None."""
    return C.c_void_p(0 if t is None else t.data_ptr())
