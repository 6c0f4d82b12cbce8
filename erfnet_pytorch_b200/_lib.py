"""ctypes binding of liberfnet_b200.so (include/erfnet_b200.h).  There is NO fallback: if the library is missing
or a call fails, a RuntimeError is raised (the reference surfaces shape errors as torch RuntimeError too)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ERF_TIMELINE=1 (tools/timeline_*.py, tools/trace_step_conv1d.py) loads the build whose kernels carry the per-CTA phase stamps; the
# product library has them compiled out (measured: the stamps' loads of the table pointer cost ~2 % of a training step, ~3 % of a forward)
LIB_PATH = os.path.join(_HERE, "lib", "liberfnet_b200_trace.so" if os.environ.get("ERF_TIMELINE", "0") == "1" else "liberfnet_b200.so")

ERF_F32, ERF_BF16 = 0, 1


class ConvGeom(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ("N", "H", "W", "Cin", "in_ld", "Ho", "Wo", "Cout", "out_ld", "out_coff",
                 "KH", "KW", "stride", "pad_h", "pad_w", "dil_h", "dil_w", "transposed", "w_ld")]


class ConvEpilogue(C.Structure):
    _fields_ = [("bias", C.c_void_p), ("relu", C.c_int), ("mask", C.c_void_p), ("mask_ld", C.c_int),
                ("add", C.c_void_p), ("add_ld", C.c_int), ("stats", C.c_void_p), ("stats_c", C.c_int)]


class Im2colDesc(C.Structure):
    _fields_ = ([("in_", C.c_void_p)] + [(n, C.c_int) for n in ("H", "W", "in_ld", "Cin", "EN", "EH", "EW", "in_step", "ntaps")] +
                [("tap_dy", C.c_int * 9), ("tap_dx", C.c_int * 9), ("w_img", C.c_void_p), ("bias", C.c_void_p), ("out", C.c_void_p)] +
                [(n, C.c_int) for n in ("Ho", "Wo", "out_ld", "out_c0", "out_step", "out_oy", "out_ox", "NP", "n_valid", "relu", "units",
                                        "in_step_x", "out_step_x", "bias_period")] +
                [("nchw_out", C.c_void_p), ("nchw_C", C.c_int)])


class WgradIm2colDesc(C.Structure):
    _fields_ = ([("g", C.c_void_p)] + [(n, C.c_int) for n in ("GH", "GW", "g_ld", "Cg", "CgP", "g_copy_bytes", "g_step")] +
                [("d", C.c_void_p)] + [(n, C.c_int) for n in ("EN", "EH", "EW", "d_ld", "d_c0", "d_c_base", "Cd", "NP", "ntaps")] +
                [("tap_dy", C.c_int * 9), ("tap_dx", C.c_int * 9), ("tap_off", C.c_int * 9), ("KHW", C.c_int), ("tail_bytes", C.c_int),
                 ("g_step_x", C.c_int), ("fold", C.c_int), ("fold_c", C.c_int)])


class WgradJob(C.Structure):
    _fields_ = [("x", C.c_void_p), ("dy", C.c_void_p), ("dw", C.c_void_p), ("horizontal", C.c_int), ("dil", C.c_int)]


class Im2colTaps(C.Structure):
    _fields_ = [("ntaps", C.c_int), ("off", C.c_longlong * 9), ("n_split", C.c_int), ("off2", C.c_longlong * 9)]


_vp, _i, _ll, _f, _d = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double
_SIGNATURES = {
    "erf_abi_version": [],
    "erf_set_launch_cap": [_i, _i],
    "erf_conv_fwd": [_vp, _vp, _vp, C.POINTER(ConvGeom), C.POINTER(ConvEpilogue), _i, _vp],
    "erf_conv_wgrad": [_vp, _vp, _vp, _vp, C.POINTER(ConvGeom), _i, _vp],
    "erf_conv_wgrad_torch": [_vp, _vp, _vp, _vp, C.POINTER(ConvGeom), _i, _i, _vp],
    "erf_stem_fwd": [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp],
    "erf_stem_fwd_nchw": [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp],
    "erf_stem_wgrad": [_vp, _vp, _i, _i, _i, _vp, _vp, _i, _vp],
    "erf_maxpool2x2_fwd": [_vp, _i, _i, _i, _i, _i, _vp, _i, _i, _vp, _vp, _i, _vp, _i, _i, _vp],
    "erf_maxpool2x2_bwd": [_vp, _i, _i, _i, _i, _i, _vp, _i, _i, _vp, _i, _i, _vp],
    "erf_nchw_f32_to_nhwc": [_vp, _i, _i, _i, _i, _vp, _i, _i, _vp],
    "erf_nhwc_to_nchw_f32": [_vp, _i, _i, _i, _i, _i, _vp, _i, _vp],
    "erf_affine_act": [_vp, _i, _vp, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "erf_channel_stats": [_vp, _i, _i, _ll, _vp, _i, _vp],
    "erf_relu_bwd": [_vp, _i, _vp, _i, _vp, _i, _ll, _i, _i, _vp],
    "erf_bn_finalize": [_vp, _d, _vp, _vp, _vp, _vp, _f, _f, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "erf_bn_eval_affine": [_vp, _vp, _vp, _vp, _f, _i, _vp, _vp, _vp],
    "erf_bn_bwd_reduce": [_vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _vp, _i, _vp],
    "erf_bn_bwd_apply": [_vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _vp],
    "erf_bn_fused_supported": [_i],
    "erf_bn_fwd_apply": [_vp, _i, _vp, _d, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp],
    "erf_bn_bwd_fused": [_vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _vp],
    "erf_bn_bwd_fused_spread": [_vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _vp],
    "erf_ce_weight_sum": [_vp, _ll, _vp, _i, _i, _vp, _vp],
    "erf_ce_fwd_bwd": [_vp, _i, _i, _vp, _ll, _vp, _i, _i, _vp, _vp, _vp, _i, _vp],
    "erf_ce_fwd_bwd_colsum_supported": [_i, _i, _i],
    "erf_ce_fwd_bwd_colsum": [_vp, _i, _vp, _ll, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "erf_confusion_hist": [_vp, _vp, _i, _ll, _i, _vp, _vp],
    "erf_iou_counts": [_vp, _i, _i, _i, _vp, _vp, _vp, _vp],
    "erf_argmax": [_vp, _i, _i, _ll, _i, _vp, _i, _i, _vp],
    "erf_masked_loss_fwd": [_i, _vp, _vp, _vp, _vp, _ll, _ll, _i, _f, _vp, _vp],
    "erf_masked_loss_bwd": [_i, _vp, _vp, _vp, _vp, _ll, _ll, _i, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "erf_softmax_nhwc_fwd": [_vp, _i, _vp, _i, _ll, _i, _vp, _i, _vp],
    "erf_softmax_nhwc_bwd": [_vp, _vp, _vp, _i, _ll, _i, _vp, _i, _i, _vp],
    "erf_adam_step": [_vp, _vp, _vp, _vp, _ll, _f, _f, _f, _f, _f, _i, _f, _vp],
    "erf_adam_step_dev": [_vp, _vp, _vp, _vp, _ll, _f, _f, _f, _f, _f, _vp, _f, _vp],
    "erf_adam_step_graph": [_vp, _vp, _vp, _vp, _ll, _vp, _f, _f, _f, _vp, _f, _vp],
    "erf_conv1d_tc_supported": [_i, _i, _i],
    "erf_conv1d_tc": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp],
    "erf_conv1d_pair_tc_supported": [_i, _i, _i],
    "erf_conv1d_pair_tc": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "erf_pack_conv1d_tc": [_vp, _i, _vp, _vp, _vp],
    "erf_pack_conv1d_tc_batched": [_vp, _i, _i, _vp],
    "erf_conv1d_wgrad_tc_workspace": [_i],
    "erf_conv1d_wgrad_tc": [_vp, _vp, _vp, _vp, C.c_size_t, _i, _i, _i, _i, _i, _i, _vp],
    "erf_conv1d_wgrad_group_tc_workspace": [_i],
    "erf_conv1d_wgrad_group_tc": [C.POINTER(WgradJob), _i, _vp, C.c_size_t, _i, _i, _i, _i, _vp],
    "erf_conv1d_wgrad_tc_fused": [_vp, _vp, _vp, _vp, C.c_size_t, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "erf_channel_sum": [_vp, _i, _i, _i, _ll, _vp, _i, _vp],
    "erf_conv_im2col_tc": [C.POINTER(Im2colDesc), _vp],
    "erf_pack_im2col_weights": [_vp, _ll, _ll, _i, _i, _i, C.POINTER(Im2colTaps), _i, _i, _vp, _vp],
    "erf_conv_wgrad_im2col_tc_workspace": [],
    "erf_conv_wgrad_im2col_tc": [C.POINTER(WgradIm2colDesc), _vp, _vp, C.c_size_t, _vp],
    "erf_avgpool_fwd": [_vp, _i, _i, _i, _i, _i, _i, _vp, _i, _vp],
    "erf_avgpool_bwd": [_vp, _i, _i, _i, _i, _i, _i, _vp, _i, _vp],
    "erf_linear_fwd": [_vp, _vp, _vp, _vp, _i, _i, _i, _vp],
    "erf_linear_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp],
    "erf_ce_rows_fwd": [_vp, _vp, _i, _i, _ll, _vp, _vp, _vp, _vp],
    "erf_ce_rows_bwd": [_vp, _vp, _i, _i, _ll, _vp, _vp, _vp, _vp, _vp],
    "erf_sgd_step": [_vp, _vp, _vp, _ll, _f, _f, _f, _i, _f, _vp],
    "erf_sgd_step_graph": [_vp, _vp, _vp, _ll, _vp, _f, _f, _vp],
    "erf_ema_update": [_vp, _vp, _ll, _f, _vp],
    "erf_ingest_u8": [_vp, _ll, _i, _vp, _vp, _vp, _i, _vp],
    "erf_label_lut": [_vp, _i, _ll, _vp, _i, _vp, _vp],
    "erf_colorize": [_vp, _i, _ll, _ll, _vp, _i, _vp, _vp],
    "erf_umma_probe": [_vp, _i, _vp, _i, _i, _i, _i, _vp, _i, _vp, _vp],
}

_lib = None


def exported_symbols():
    """Every symbol include/erfnet_b200.h declares (used by the CPU test that checks the library exports them)."""
    return ["erf_last_error_string"] + list(_SIGNATURES)


_TOOLING = {"erf_set_launch_cap"}   # entry points only experiments use: an A/B against an older build may lack them


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m erfnet_pytorch_b200.build` "
                               "(there is no CPU / eager fallback)")
        L = C.CDLL(LIB_PATH)
        L.erf_last_error_string.restype = C.c_char_p
        L.erf_last_error_string.argtypes = []
        for name, args in _SIGNATURES.items():
            if name in _TOOLING and not hasattr(L, name):
                continue                       # an older build of the library under tools/ab_libs.sh
            fn = getattr(L, name)
            fn.restype = C.c_size_t if name.endswith("_workspace") else C.c_int
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise RuntimeError(f"liberfnet_b200: {lib().erf_last_error_string().decode()} (code {rc})")
