"""ctypes binding of libfabric_b200.so (the C ABI in include/fabric_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no fallback: if the
library is missing, or the device is not sm_100, every op raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libfabric_b200.so")

FB_OK = 0


class FabricB200Error(RuntimeError):
    pass


class ConvTuning(C.Structure):
    _fields_ = [("n_tile", C.c_int), ("halo", C.c_int), ("a_stages", C.c_int), ("b_stages", C.c_int),
                ("b_resident", C.c_int), ("grid", C.c_int), ("ctas", C.c_int), ("epi_warps", C.c_int), ("occupancy", C.c_int)]


class Conv3x3Desc(C.Structure):
    _fields_ = [
        ("G", C.c_int), ("B", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("Cin", C.c_int), ("Cout", C.c_int), ("relu", C.c_int), ("store_main", C.c_int),
        ("x", C.c_void_p), ("w", C.c_void_p), ("y", C.c_void_p),
        ("scale", C.c_void_p), ("shift", C.c_void_p),
        ("pool_out", C.c_void_p), ("stats_ws", C.c_void_p),
        ("head_w", C.c_void_p), ("head_b", C.c_void_p), ("head_out", C.c_void_p),
        ("prod_out", C.c_void_p), ("prod_channels", C.c_int), ("shift_in_acc", C.c_int),
        ("tune", ConvTuning),
        ("bnbwd_z", C.c_void_p), ("bnbwd_coef", C.c_void_p),
    ]


class ConvPlan(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("n_tile", "ck", "halo", "grid", "smem_bytes", "ctas", "epi_warps", "a_stages",
                                       "b_stages", "b_resident", "out_bufs", "total_units", "pool_tma", "prod_tma", "ctas_per_sm",
                                       "reg_stats")]


class WgradDesc(C.Structure):
    _fields_ = [("G", C.c_int), ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Ca", C.c_int), ("Cb", C.c_int),
                ("p", C.c_void_p), ("q", C.c_void_p), ("ws", C.c_void_p), ("splits", C.c_int), ("wide", C.c_int)]


class WgradPlan(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("form", "grid", "items", "splits", "stages", "smem_bytes", "tiles_total")]


# name -> (restype, argtypes); every symbol declared in include/fabric_b200.h
_vp, _i, _f, _i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64
SIGNATURES = {
    "fabric_b200_version": (_i, []),
    "fabric_b200_last_error": (C.c_char_p, []),
    "fabric_b200_sm_count": (_i, []),
    "fabric_b200_pack_nchw_f32_to_nhwc_bf16": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "fabric_b200_pack_nchw_u16_to_nhwc_bf16": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "fabric_b200_pack_nchw_aug": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "fabric_b200_augment_labels": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "fabric_b200_unpack_nhwc_bf16_to_nchw_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "fabric_b200_pack_conv3x3_weight": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "fabric_b200_pack_conv3x3_weight_scaled": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "fabric_b200_conv3x3": (_i, [C.POINTER(Conv3x3Desc), _vp]),
    "fabric_b200_conv3x3_plan": (_i, [C.POINTER(Conv3x3Desc), _i, _i, C.POINTER(ConvPlan)]),
    "fabric_b200_conv3x3_grid": (_i, [C.POINTER(Conv3x3Desc)]),
    "fabric_b200_conv3x3_stats_ws_floats": (C.c_int64, [C.POINTER(Conv3x3Desc)]),
    "fabric_b200_bn_fold_eval": (_i, [_vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _i, _vp]),
    "fabric_b200_build_up_input": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "fabric_b200_outconv": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "fabric_b200_bn_finalize": (_i, [_vp, _i, _i, _i, _i, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp]),
    "fabric_b200_bn_apply_relu": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "fabric_b200_bn_apply_relu_head": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "fabric_b200_bn_head_bwd_ws_floats": (_i64, []),
    "fabric_b200_bn_head_bwd": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i,
                                     _f, _f, _vp]),
    "fabric_b200_seg_loss_ws_floats": (_i64, [_i, _i, _i]),
    "fabric_b200_seg_loss_fwd_bwd": (_i, [_i, _f, _f, _f, _f, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "fabric_b200_seg_loss_sums_offset": (_i64, [_i, _i, _i]),
    "fabric_b200_seg_loss_phase": (_i, [_i, _i, _f, _f, _f, _f, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _f, _vp]),
    "fabric_b200_outconv_bwd_ws_floats": (_i64, [_i]),
    "fabric_b200_outconv_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "fabric_b200_bn_bwd_ws_floats": (_i64, [_i, _i]),
    "fabric_b200_bn_relu_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                     _i, _i, _i, _i, _i, _vp]),
    "fabric_b200_bn_bwd_partial_floats": (_i64, [_i, _i]),
    "fabric_b200_bn_bwd_from_partials": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i,
                                              _f, _f, _vp]),
    "fabric_b200_bn_relu_bwd_phase": (_i, [_i, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                           _i, _i, _i, _i, _i, _f, _f, _vp]),
    "fabric_b200_up_input_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "fabric_b200_conv3x3_wgrad_plan": (_i, [C.POINTER(WgradDesc), _i, _i, C.POINTER(WgradPlan)]),
    "fabric_b200_conv3x3_wgrad_ws_floats": (_i64, [C.POINTER(WgradDesc)]),
    "fabric_b200_conv3x3_wgrad_splits": (_i, [C.POINTER(WgradDesc)]),
    "fabric_b200_conv3x3_wgrad": (_i, [C.POINTER(WgradDesc), _vp]),
    "fabric_b200_wgrad_reduce": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "fabric_b200_wgrad_reduce_swapped": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "fabric_b200_gather_tiles": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "fabric_b200_argmax_metrics": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "fabric_b200_scatter_tiles": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "fabric_b200_sgd_step": (_i, [_vp, _i, _f, _f, _vp]),
    "fabric_b200_train_step_update": (_i, [_vp, _i, _f, _f, _f, _vp]),
}

_lib = None


def load():
    """Load the shared library (once) and declare all prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FabricB200Error(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). fabric_b200 has no CPU or eager fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc < 0:
        msg = load().fabric_b200_last_error().decode("utf-8", "replace")
        raise FabricB200Error(f"{what or 'fabric_b200'} failed ({rc}): {msg}")
    return rc
