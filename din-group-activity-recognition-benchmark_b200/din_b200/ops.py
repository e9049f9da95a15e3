"""torch-tensor wrappers over the C ABI (pointer extraction, shape checks, output allocation).

Every function enqueues on torch's current CUDA stream and returns immediately.  Nothing here
computes: the arithmetic is in libdin_sm100.so.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import DinConvDesc, check


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _need(t, dtype, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise _lib.DinError(f"{name}: expected a contiguous CUDA {dtype} tensor, got "
                            f"{getattr(t, 'dtype', type(t))} on {getattr(t, 'device', '?')}")


def pack_conv_weight(w_oihw, scale=None, c_in_padded=None):
    """OIHW fp32 -> packed fp16 [c_out][kh][kw][c_in_padded] (optionally folding a per-channel scale)."""
    _need(w_oihw, torch.float32, "w_oihw")
    co, ci, kh, kw = w_oihw.shape
    cip = ci if c_in_padded is None else c_in_padded
    if scale is not None:
        _need(scale, torch.float32, "scale")
    out = torch.empty((co, kh, kw, cip), dtype=torch.float16, device=w_oihw.device)
    check(_lib.load().din_pack_conv_weight_f16(_p(w_oihw), _p(scale), _p(out), co, ci, cip, kh, kw, _stream()),
          "din_pack_conv_weight_f16")
    return out


def conv2d_nhwc(x, w_packed, bias=None, *, stride=1, pad=(0, 0), relu=False, residual=None, out=None,
                out_f32=False, c_in=None, x_c_offset=0, y_c_offset=0, c_out=None):
    """x: [n,h,w,Cx] fp16 NHWC (the conv reads channels [x_c_offset, x_c_offset+c_in)).
    w_packed: [c_out, kh, kw, c_in] fp16.  Returns / fills `out` [n,oh,ow,Cy] at channel offset y_c_offset."""
    _need(x, torch.float16, "x")
    _need(w_packed, torch.float16, "w_packed")
    n, h, w, cx = x.shape
    co, kh, kw, ci = w_packed.shape
    if c_in is None:
        c_in = ci
    assert c_in == ci, (c_in, ci)
    ph, pw = pad
    oh = (h + 2 * ph - kh) // stride + 1
    ow = (w + 2 * pw - kw) // stride + 1
    if out is None:
        out = torch.empty((n, oh, ow, co), dtype=torch.float32 if out_f32 else torch.float16, device=x.device)
    _need(out, torch.float32 if out_f32 else torch.float16, "out")
    assert out.shape[:3] == (n, oh, ow), (out.shape, (n, oh, ow))
    cy = out.shape[3]
    d = DinConvDesc(n=n, h=h, w=w, c_in=c_in, x_c_stride=cx, c_out=co, y_c_stride=cy, kh=kh, kw=kw,
                    stride=stride, pad_h=ph, pad_w=pw, relu=int(relu), out_f32=int(out_f32))
    esz_y = 4 if out_f32 else 2
    xp = C.c_void_p(x.data_ptr() + 2 * x_c_offset)
    yp = C.c_void_p(out.data_ptr() + esz_y * y_c_offset)
    rp = C.c_void_p(0)
    if residual is not None:
        _need(residual, torch.float16, "residual")
        assert residual.shape == out.shape
        rp = C.c_void_p(residual.data_ptr() + 2 * y_c_offset)
    if bias is not None:
        _need(bias, torch.float32, "bias")
    check(_lib.load().din_conv2d_nhwc_f16(C.byref(d), xp, _p(w_packed), _p(bias), rp, yp, _stream()),
          "din_conv2d_nhwc_f16")
    return out


def stem_conv(x_nchw, w_oihw, bias, *, stride=1, pad=0, relu=True, prep=True):
    """Raw fp32 NCHW images (0..255) -> prep_images -> conv(+bias,+ReLU) -> NHWC fp16."""
    _need(x_nchw, torch.float32, "x_nchw")
    _need(w_oihw, torch.float32, "w_oihw")
    n, c, h, w = x_nchw.shape
    assert c == 3
    co, ci, kh, kw = w_oihw.shape
    assert ci == 3
    oh = (h + 2 * pad - kh) // stride + 1
    ow = (w + 2 * pad - kw) // stride + 1
    y = torch.empty((n, oh, ow, co), dtype=torch.float16, device=x_nchw.device)
    if bias is not None:
        _need(bias, torch.float32, "bias")
    check(_lib.load().din_stem_conv_nchw_f32(_p(x_nchw), _p(w_oihw), _p(bias), _p(y), n, h, w, co, kh, kw,
                                             stride, pad, int(relu), int(prep), _stream()),
          "din_stem_conv_nchw_f32")
    return y


def maxpool2d_nhwc(x, k, stride, pad=0):
    _need(x, torch.float16, "x")
    n, h, w, c = x.shape
    oh = (h + 2 * pad - k) // stride + 1
    ow = (w + 2 * pad - k) // stride + 1
    y = torch.empty((n, oh, ow, c), dtype=torch.float16, device=x.device)
    check(_lib.load().din_maxpool2d_nhwc_f16(_p(x), _p(y), n, h, w, c, k, stride, pad, _stream()),
          "din_maxpool2d_nhwc_f16")
    return y
