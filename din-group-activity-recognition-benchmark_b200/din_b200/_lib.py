"""ctypes binding of the C ABI in include/din_sm100.h (one prototype per exported symbol)."""
import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libdin_sm100.so")

DIN_OK = 0


class DinError(RuntimeError):
    pass


class DinConvDesc(C.Structure):
    _fields_ = [(name, C.c_int32) for name in (
        "n", "h", "w", "c_in", "x_c_stride", "c_out", "y_c_stride", "kh", "kw", "stride",
        "pad_h", "pad_w", "relu", "out_f32", "pool2", "w_split")]


class DinPackJob(C.Structure):
    _fields_ = [("w", C.c_void_p), ("scale", C.c_void_p), ("out", C.c_void_p)] + [(name, C.c_int32) for name in (
        "rows", "cols", "cols_padded", "kh", "kw", "split", "transposed", "reserved")]


class DinConvBranchOut(C.Structure):
    _fields_ = [(name, C.c_int32) for name in ("split_col", "y2_c_stride", "norelu_lo", "norelu_hi")]


class DinFlatJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst_offset", C.c_longlong), ("numel", C.c_longlong)]


_vp, _i, _fp, _ll = C.c_void_p, C.c_int, C.c_void_p, C.c_longlong  # float* is passed as a raw address

# name -> (restype, argtypes).  tests/test_abi_cpu.py checks this table against include/din_sm100.h.
PROTOTYPES = {
    "din_abi_version": (C.c_int, []),
    "din_last_error_string": (C.c_char_p, []),
    "din_device_sm_count": (C.c_int, []),
    "din_debug_word": (C.c_uint, [_i]),
    "din_tmap_cache_stats": (C.c_int, [C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]),
    "din_stem_conv_nchw_f32": (C.c_int, [_fp, _fp, _fp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "din_stem_conv_nhwc_u8": (C.c_int, [_vp, _fp, _fp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "din_conv2d_nhwc_f16": (C.c_int, [C.POINTER(DinConvDesc), _vp, _vp, _fp, _vp, _vp, _vp]),
    "din_conv2d_branches_nhwc_f16": (C.c_int, [C.POINTER(DinConvDesc), C.POINTER(DinConvBranchOut), _vp, _vp, _fp, _vp,
                                               _vp, _vp]),
    "din_conv3x3_stem_pair_nhwc_f16": (C.c_int, [_vp, _i, _fp, _fp, _vp, _fp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "din_conv2d_relu_bwd_nhwc_f16": (C.c_int, [C.POINTER(DinConvDesc), _vp, _vp, _vp, _vp, _vp]),
    "din_pack_conv_weight_f16": (C.c_int, [_fp, _fp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "din_bn_fold_grads_f32": (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_float, _fp, _i, _ll, _vp]),
    "din_bn_stats": (C.c_int, [_vp, _i, _ll, _i, _vp, _vp, _vp]),
    "din_bn_finalize_f32": (C.c_int, [_vp, _vp, _ll, _fp, _fp, C.c_float, C.c_float, _fp, _fp, _fp, _fp, _fp, _fp, _i, _vp]),
    "din_bn_apply": (C.c_int, [_vp, _i, _fp, _fp, _vp, _vp, _ll, _i, _i, _vp]),
    "din_bn_bwd": (C.c_int, [_vp, _vp, _i, _fp, _fp, _fp, _fp, _vp, _fp, _fp, _fp, _ll, _i, _vp]),
    "din_pack_conv_weights_f16": (C.c_int, [C.POINTER(DinPackJob), _i, _vp]),
    "din_stem7x7_pool_nhwc_f16": (C.c_int, [_vp, _i, _fp, _fp, _vp, _i, _i, _i, _i, _vp]),
    "din_maxpool2d_nhwc_f16": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "din_avgpool2d_nhwc_f16": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "din_avgpool3_bias_relu_nhwc_f16": (C.c_int, [_vp, _vp, _fp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "din_upsample_bilinear_nhwc_f16": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "din_roi_align_nhwc_f16": (C.c_int, [_vp, _fp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "din_roi_align_nhwc_f16_f32out": (C.c_int, [_vp, _fp, _vp, _fp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "din_group_layernorm_f32": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _ll, _ll, _i, _ll, _i,
                                          C.c_float, _i, _vp, _vp]),
    "din_linear_f32": (C.c_int, [_fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _vp]),
    "din_dynamic_infer_f32": (C.c_int, [_fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _i, _i, _fp, C.c_float,
                                        _i, _vp, _vp]),
    "din_readout_f32": (C.c_int, [_fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _vp, _vp]),
    "din_ce_metrics_f32": (C.c_int, [_fp, _vp, _fp, C.c_float, _fp, _vp, _vp, _vp, _fp, _i, _i, _vp]),
    "din_mean_axis_f32": (C.c_int, [_fp, _fp, _i, _i, _i, _vp]),
    "din_conv2d_wgrad_nhwc_f16": (C.c_int, [_vp, _vp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "din_roi_align_bwd_f32": (C.c_int, [_fp, _fp, _vp, _fp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "din_grad_to_f16": (C.c_int, [_fp, _vp, _fp, _ll, C.c_float, _vp]),
    "din_relu_pool_bwd_nhwc_f16": (C.c_int, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "din_stem_wgrad": (C.c_int, [_vp, _i, _vp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "din_scatter2_nhwc_f16": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "din_add_f16": (C.c_int, [_vp, _vp, _vp, _ll, _vp]),
    "din_maxpool3s2_relu_bwd_nhwc_f16": (C.c_int, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "din_resize_bilinear_u8": (C.c_int, [_vp, _i, _i, _i, _vp, _i, _i, _vp, _vp]),
    "din_jpeg_backend": (C.c_int, []),
    "din_jpeg_image_info": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "din_jpeg_decode_resize_u8": (C.c_int, [C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), _i, _vp, _i, _i, _vp, C.c_size_t,
                                            _i, _vp]),
    "din_maxpool3s2_bwd_nhwc_f16": (C.c_int, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "din_relu_bwd_slice_nhwc_f16": (C.c_int, [_vp, _vp, _vp, _ll, _i, _i, _i, _i, _vp]),
    "din_upsample_bilinear_bwd_nhwc_f16": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "din_colsum_nhwc_f16": (C.c_int, [_vp, _fp, _ll, _i, _i, _fp, _vp]),
    "din_bn_gamma_grad_f16": (C.c_int, [_vp, _vp, _vp, _fp, _fp, _fp, _fp, _ll, _i, _vp]),
    "din_scale_rows_f32": (C.c_int, [_fp, _fp, _ll, _ll, _vp]),
    "din_gemm_f32": (C.c_int, [_fp, _ll, _ll, _vp, _i, _ll, _ll, _fp, _ll, _i, _i, _i, C.c_float, _i, _vp]),
    "din_colsum_f32": (C.c_int, [_fp, _fp, _i, _i, _ll, _vp]),
    "din_scale_mask_f32": (C.c_int, [_fp, _vp, C.c_float, _fp, _ll, _vp]),
    "din_relu_bwd_f32": (C.c_int, [_fp, _fp, _fp, _ll, _vp]),
    "din_readout_bwd_f32": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _vp, _vp]),
    "din_group_layernorm_bwd_f32": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _ll, _ll, _i,
                                              _ll, _i, C.c_float, _i, _i, _vp, _vp]),
    "din_context_attention_f32": (C.c_int, [_fp, _fp, _fp, _fp, _i, _i, _i, _i, _vp]),
    "din_context_attention_bwd_f32": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _vp]),
    "din_pack_flat_f32": (C.c_int, [C.POINTER(DinFlatJob), _i, _fp, C.c_float, _vp]),
    "din_dynamic_infer_bwd_ws_floats": (C.c_longlong, [_i, _i, _i, _i, _i, _i, _i]),
    "din_dynamic_infer_bwd_f32": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i,
                                            _i, _i, _fp, C.c_float, _vp, _vp]),
}

_lock = threading.Lock()
_lib = None


def load():
    """Load libdin_sm100.so (built in-tree by __graft_entry__.build()). Fails loudly if absent."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise DinError(
                    f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(there is no CPU fallback for the DIN hot path)")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in PROTOTYPES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(rc, what=""):
    if rc != DIN_OK:
        msg = load().din_last_error_string().decode("utf-8", "replace")
        raise DinError(f"{what or 'libdin_sm100'} failed (code {rc}): {msg}")
    return rc
