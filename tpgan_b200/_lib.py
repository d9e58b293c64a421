"""ctypes binding of the C ABI in include/tpgan_b200.h (libtpgan_b200.so).

The product path has no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libtpgan_b200.so")

CONV_FWD, CONV_DGRAD, DECONV_FWD, DECONV_DGRAD = 0, 1, 2, 3
EPI_LINEAR, EPI_LEAKY, EPI_MASK = 0, 1, 2
DTYPE_TF32, DTYPE_BF16 = 0, 1


class View(C.Structure):
    """tpgan_view: NHWC fp32 view with unit channel stride (strides in elements)."""

    _fields_ = [("ptr", C.c_void_p), ("sn", C.c_int64), ("sh", C.c_int64), ("sw", C.c_int64),
                ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32)]


class ConvArgs(C.Structure):
    _fields_ = [("kind", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
                ("in_", View), ("out", View), ("w_packed", C.c_void_p), ("w_rows_pad", C.c_int32),
                ("w_k_pad", C.c_int32), ("bias", C.c_void_p), ("add1", View), ("add2", View), ("mask", View),
                ("slopes", C.c_void_p), ("slope", C.c_float), ("epilogue", C.c_int32), ("round_tf32", C.c_int32),
                ("dtype", C.c_int32), ("out16", View), ("in_lo", View), ("w_lo_packed", C.c_void_p)]


class WgradArgs(C.Structure):
    _fields_ = [("kind", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
                ("x", View), ("dy", View), ("dw_packed", C.c_void_p), ("w_rows_pad", C.c_int32), ("w_k_pad", C.c_int32),
                ("accumulate", C.c_int32), ("dtype", C.c_int32)]


class CastJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("rows", C.c_int64), ("k_pad", C.c_int32), ("k_pad16", C.c_int32),
                ("block_begin", C.c_int32), ("pad_", C.c_int32)]


class BiasJob(C.Structure):
    _fields_ = [("dy", C.c_void_p), ("db", C.c_void_p), ("npix", C.c_int64), ("sw", C.c_int64), ("c", C.c_int32),
                ("block_begin", C.c_int32), ("pix_blocks", C.c_int32), ("cgroups", C.c_int32), ("lanes", C.c_int32),
                ("pad_", C.c_int32)]


class PackJob(C.Structure):
    _fields_ = [("ref_c", C.c_void_p), ("ref", C.c_void_p), ("packed", C.c_void_p), ("row_map", C.c_void_p),
                ("k_map", C.c_void_p), ("rs", C.c_int64), ("taps", C.c_int32), ("rows", C.c_int32), ("k", C.c_int32),
                ("rows_pad", C.c_int32), ("k_pad", C.c_int32), ("row_len", C.c_int32), ("flag", C.c_int32),
                ("block_begin", C.c_int32)]


class TransposeJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("taps", C.c_int32), ("rows", C.c_int32), ("k", C.c_int32),
                ("rows_src_pad", C.c_int32), ("k_src_pad", C.c_int32), ("rows_dst_pad", C.c_int32),
                ("k_dst_pad", C.c_int32), ("block_begin", C.c_int32), ("tiles_k", C.c_int32), ("tiles_r", C.c_int32)]


NULL_VIEW = View(None, 0, 0, 0, 0, 0, 0, 0)

_lib: Optional[C.CDLL] = None

# name -> (restype, argtypes); every symbol include/tpgan_b200.h declares
_VP, _I32, _I64, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
SYMBOLS = {
    "tpgan_conv2d": (C.c_int, [C.POINTER(ConvArgs), _I32, _VP]),
    "tpgan_conv2d_wgrad": (C.c_int, [C.POINTER(WgradArgs), _I32, _VP]),
    "tpgan_pack_weights": (C.c_int, [_VP, _VP, _I32, _I32, _I32, _I32, _I32, _I64, _I64, _VP, _VP, _I32, _VP]),
    "tpgan_unpack_weights": (C.c_int, [_VP, _VP, _I32, _I32, _I32, _I32, _I32, _I64, _I64, _VP, _VP, _I32, _VP]),
    "tpgan_transpose_packed": (C.c_int, [_VP, _VP, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _VP]),
    "tpgan_bias_grad_multi": (C.c_int, [_VP, _I32, _I32, _VP]),
    "tpgan_cast_packed_multi": (C.c_int, [_VP, _I32, _I32, _VP]),
    "tpgan_cast_bf16": (C.c_int, [View, View, _VP]),
    "tpgan_pack_multi": (C.c_int, [_VP, _I32, _I32, _I32, _I32, _VP]),
    "tpgan_transpose_multi": (C.c_int, [_VP, _I32, _I32, _VP]),
    "tpgan_nchw_to_nhwc": (C.c_int, [_VP, View, _I32, _VP]),
    "tpgan_nhwc_to_nchw": (C.c_int, [View, _VP, _VP]),
    "tpgan_act_backward": (C.c_int, [View, View, View, _VP, _F, _VP]),
    "tpgan_view_copy": (C.c_int, [View, View, _I32, _VP]),
    "tpgan_bias_grad": (C.c_int, [View, _VP, _I32, _VP]),
    "tpgan_reflect_pad": (C.c_int, [View, View, _I32, _I32, _VP]),
    "tpgan_reflect_pad_backward": (C.c_int, [View, View, _I32, _I32, _I32, _VP]),
    "tpgan_patch_crop": (C.c_int, [View, _VP, View, View, View, View, _VP, _F, _VP]),
    "tpgan_local_fuse": (C.c_int, [View, View, View, View, View, _VP, _VP]),
    "tpgan_local_fuse_backward": (C.c_int, [View, _VP, View, View, View, View, _I32, _VP]),
    "tpgan_maxpool3s2": (C.c_int, [View, View, _VP, _VP]),
    "tpgan_maxpool3s2_backward": (C.c_int, [View, _VP, View, _I32, _VP]),
    "tpgan_avgpool": (C.c_int, [View, View, _VP]),
    "tpgan_avgpool_backward": (C.c_int, [View, View, _I32, _VP]),
    "tpgan_image_losses": (C.c_int, [View, View, View, View, View, _VP, _VP, _VP]),
    "tpgan_l1_loss": (C.c_int, [View, View, View, _F, _VP, _VP]),
    "tpgan_maxout2": (C.c_int, [_VP, _VP, _I32, _I32, _VP]),
    "tpgan_maxout2_backward": (C.c_int, [_VP, _VP, _VP, _I32, _I32, _VP]),
    "tpgan_adam_step": (C.c_int, [_VP, _VP, _VP, _VP, _I64, _F, _F, _F, _F, _F, _I32, _F, _VP]),
    "tpgan_adam_step_dev": (C.c_int, [_VP, _VP, _VP, _VP, _I64, _F, _F, _F, _F, _F, _VP, _F, _VP]),
    "tpgan_adam_slice_dev": (C.c_int, [_VP, _VP, _VP, _VP, _I64, _F, _F, _F, _F, _F, _VP, _F, _I32, _VP]),
    "tpgan_sample_sqnorm": (C.c_int, [View, _VP, _VP]),
    "tpgan_sample_scale": (C.c_int, [View, _VP, View, _VP]),
    "tpgan_gp_coeff": (C.c_int, [_VP, _VP, _I32, _F, _VP, _VP]),
    "tpgan_lerp": (C.c_int, [View, View, _VP, View, _VP]),
    "tpgan_mul": (C.c_int, [View, View, View, _VP]),
    "tpgan_fill": (C.c_int, [View, _F, _VP]),
    "tpgan_split_tf32": (C.c_int, [View, View, View, _VP]),
    "tpgan_softmax_ce": (C.c_int, [_VP, _I64, _VP, _VP, _I64, _I32, _I32, _F, _VP, _VP]),
    "tpgan_dwconv3x3": (C.c_int, [View, View, _VP, _I32, _VP]),
    "tpgan_dwconv3x3_dgrad": (C.c_int, [View, View, _VP, _I32, _I32, _VP]),
    "tpgan_dwconv3x3_wgrad": (C.c_int, [View, View, _VP, _I32, _VP]),
    "tpgan_bn_forward": (C.c_int, [View, View, View, _VP, _VP, _VP, _VP, _F, _F, _I32, _I32, _F, _I32, _VP, _VP, _VP]),
    "tpgan_bn_backward": (C.c_int, [View, View, View, _VP, _I32, _I32, _I32, _I32, _VP, _VP, _VP, _VP]),
    "tpgan_rows_gather": (C.c_int, [View, _VP, _I64, _I64, _I32, _VP]),
    "tpgan_multitask_loss": (C.c_int, [_VP, _VP, _VP, _VP, _I32, _I32, _I64, _I64, _I32, _I32, _F, _F, _F, _F, _F, _F, _VP, _VP, _VP,
                                       _VP, _VP]),
    "tpgan_ssd_decode": (C.c_int, [_VP, _VP, _I32, _I32, _I64, _I64, _I32, _I32, _F, _F, _VP, _VP, _VP, _VP, _VP, _VP]),
    "tpgan_sgd_step": (C.c_int, [_VP, _VP, _VP, _I64, _VP, _F, _F, _I32, _F, _VP]),
    "tpgan_u8_to_nhwc": (C.c_int, [_VP, View, _I32, _VP]),
    "tpgan_landmarks_reduce": (C.c_int, [_VP, _I32, _I32, _VP, _I32, _F, _F, _VP, _VP]),
    "tpgan_pyramid": (C.c_int, [View, View, View, _VP]),
    "tpgan_last_error": (C.c_char_p, []),
    "tpgan_abi_version": (C.c_int, []),
    "tpgan_kernel_status": (C.c_int, []),
    "tpgan_launch_count": (C.c_int64, []),
    "tpgan_last_conv_kernel": (C.c_int, []),
    "tpgan_last_conv_pair": (C.c_int, []),
    "tpgan_set_deterministic": (C.c_int, [_I32]),
    "tpgan_get_deterministic": (C.c_int, []),
    "tpgan_set_sm_reserve": (C.c_int, [_I32]),
}


def load() -> C.CDLL:
    """Load libtpgan_b200.so (built in-tree by tpgan_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"tpgan_b200: CUDA extension not built ({LIB_PATH} missing). Run `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C tpgan_b200/csrc`. There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError here = ABI mismatch, fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().tpgan_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"tpgan_b200 {what} failed (status {rc}): {msg}")


def launch_count() -> int:
    return int(load().tpgan_launch_count())


def last_conv_pair() -> bool:
    """True when the calling thread's most recent conv launch ran over CTA pairs (cta_group::2)."""
    return bool(load().tpgan_last_conv_pair())


def last_conv_kernel() -> str:
    """Which kernel the most recent conv2d call of this thread launched."""
    return ("tapgemm", "rowconv", "rowstack", "flatconv")[int(load().tpgan_last_conv_kernel())]


def set_deterministic(on: bool) -> bool:
    """Process-wide deterministic mode (see include/tpgan_b200.h); returns the previous setting.  Plans / job tables built
    while it is on keep their deterministic geometry."""
    return bool(load().tpgan_set_deterministic(int(bool(on))))


def set_sm_reserve(sms: int) -> int:
    """SMs the persistent tensor-core kernels leave free for concurrent kernels (NCCL); returns the previous value."""
    return int(load().tpgan_set_sm_reserve(int(sms)))


def deterministic() -> bool:
    return bool(load().tpgan_get_deterministic())


def kernel_status() -> int:
    return int(load().tpgan_kernel_status())
