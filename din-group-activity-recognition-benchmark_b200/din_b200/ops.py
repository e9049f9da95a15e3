"""torch-tensor wrappers over the C ABI (pointer extraction, shape checks, output allocation).

Every function enqueues on torch's current CUDA stream and returns immediately.  Nothing here
computes: the arithmetic is in libdin_sm100.so.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import DinConvBranchOut, DinConvDesc, check


LAUNCHES = 0      # kernels launched through this module (each C-ABI compute call launches exactly one)
RECORDER = None   # optional list: bench.py's profiling pass appends (name, flops, bytes, ev_start, ev_end)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _launch:
    """Counts the launch and, when a recorder is installed, brackets it with CUDA events on the launching
    (current) stream."""

    def __init__(self, name, flops=0, nbytes=0):
        self.name, self.flops, self.nbytes = name, flops, nbytes

    def __enter__(self):
        global LAUNCHES
        LAUNCHES += 1
        if RECORDER is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if RECORDER is not None:
            self.e1.record()
            RECORDER.append((self.name, self.flops, self.nbytes, self.e0, self.e1))
        return False


def _p(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _need(t, dtype, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise _lib.DinError(f"{name}: expected a contiguous CUDA {dtype} tensor, got "
                            f"{getattr(t, 'dtype', type(t))} on {getattr(t, 'device', '?')}")


def pack_conv_weight(w_oihw, scale=None, c_in_padded=None, split=1):
    """OIHW fp32 -> packed fp16 [c_out][kh][kw][c_in_padded] (optionally folding a per-channel scale);
    split=2 -> [c_out][2][kh][kw][c_in_padded] (hi, lo parts)."""
    _need(w_oihw, torch.float32, "w_oihw")
    co, ci, kh, kw = w_oihw.shape
    cip = (ci + 63) // 64 * 64 if c_in_padded is None else c_in_padded
    if scale is not None:
        _need(scale, torch.float32, "scale")
    shape = (co, kh, kw, cip) if split == 1 else (co, 2, kh, kw, cip)
    out = torch.empty(shape, dtype=torch.float16, device=w_oihw.device)
    check(_lib.load().din_pack_conv_weight_f16(_p(w_oihw), _p(scale), _p(out), co, ci, cip, kh, kw, split,
                                               _stream()), "din_pack_conv_weight_f16")
    return out


def pack_conv_weights(jobs):
    """Several weights packed by one launch.  jobs: [(w_oihw fp32, scale | None, split, transposed)].
    transposed: the data-gradient filter of the convolution with forward weight w (rot180, channel axes swapped,
    scale per forward output channel) -- no permute / flip copies.  -> [packed fp16 tensors]."""
    if not jobs:
        return []
    arr = (_lib.DinPackJob * len(jobs))()
    outs = []
    for j, (w, scale, split, transposed) in zip(arr, jobs):
        _need(w, torch.float32, "w_oihw")
        a, b, kh, kw = w.shape
        rows, cols = (b, a) if transposed else (a, b)
        cp = (cols + 63) // 64 * 64
        if scale is not None:
            _need(scale, torch.float32, "scale")
            assert scale.numel() == a
        out = torch.empty((rows, kh, kw, cp) if split == 1 else (rows, 2, kh, kw, cp), dtype=torch.float16, device=w.device)
        outs.append(out)
        j.w, j.scale, j.out = w.data_ptr(), (0 if scale is None else scale.data_ptr()) or None, out.data_ptr()
        j.rows, j.cols, j.cols_padded, j.kh, j.kw, j.split, j.transposed = rows, cols, cp, kh, kw, split, int(transposed)
    check(_lib.load().din_pack_conv_weights_f16(arr, len(jobs), _stream()), "din_pack_conv_weights_f16")
    return outs


def conv2d_nhwc(x, w_packed, bias=None, *, stride=1, pad=(0, 0), relu=False, residual=None, out=None,
                out_f32=False, c_in=None, x_c_offset=0, y_c_offset=0, c_out=None, pool2=False, relu_mask=None):
    """x: [n,h,w,Cx] fp16 NHWC (the conv reads channels [x_c_offset, x_c_offset+c_in)).
    w_packed: [c_out, kh, kw, c_in] fp16.  Returns / fills `out` [n,oh,ow,Cy] at channel offset y_c_offset.
    relu_mask: saved ReLU output shaped like the result -> result * [relu_mask > 0] (din_conv2d_relu_bwd_nhwc_f16: a
    data-gradient convolution fused with the ReLU backward of the layer below)."""
    _need(x, torch.float16, "x")
    _need(w_packed, torch.float16, "w_packed")
    n, h, w, cx = x.shape
    split = 2 if w_packed.dim() == 5 else 1
    co, kh, kw, ci = w_packed.shape[0], w_packed.shape[-3], w_packed.shape[-2], w_packed.shape[-1]
    # ci = c_in rounded up to a multiple of 64 (zero columns)
    if c_in is None:
        c_in = min(ci, cx - x_c_offset)
    assert (c_in + 63) // 64 * 64 == ci and x_c_offset + c_in <= cx, (c_in, ci, cx, x_c_offset)
    ph, pw = pad
    oh = (h + 2 * ph - kh) // stride + 1
    ow = (w + 2 * pw - kw) // stride + 1
    yh, yw = (oh // 2, ow // 2) if pool2 else (oh, ow)
    if out is None:
        out = torch.empty((n, yh, yw, co), dtype=torch.float32 if out_f32 else torch.float16, device=x.device)
    _need(out, torch.float32 if out_f32 else torch.float16, "out")
    assert out.shape[:3] == (n, yh, yw), (out.shape, (n, yh, yw))
    cy = out.shape[3]
    assert y_c_offset + co <= cy, (y_c_offset, co, cy)
    d = DinConvDesc(n=n, h=h, w=w, c_in=c_in, x_c_stride=cx, c_out=co, y_c_stride=cy, kh=kh, kw=kw,
                    stride=stride, pad_h=ph, pad_w=pw, relu=int(relu), out_f32=int(out_f32), pool2=int(pool2),
                    w_split=split)
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
    flops = 2 * n * oh * ow * co * kh * kw * c_in       # algorithmic: the real c_in, not the K padding
    nbytes = 2 * n * h * w * c_in + 2 * co * kh * kw * ci + esz_y * n * yh * yw * co
    if relu_mask is not None:
        _need(relu_mask, torch.float16, "relu_mask")
        assert relu_mask.shape == out.shape and residual is None and bias is None and not (relu or pool2 or out_f32)
        with _launch(f"conv{kh}x{kw}s{stride}_{c_in}->{co}@{oh}x{ow}+relubwd", flops, nbytes + 2 * out.numel()):
            check(_lib.load().din_conv2d_relu_bwd_nhwc_f16(C.byref(d), xp, _p(w_packed),
                                                           C.c_void_p(relu_mask.data_ptr() + 2 * y_c_offset), yp,
                                                           _stream()), "din_conv2d_relu_bwd_nhwc_f16")
        return out
    with _launch(f"conv{kh}x{kw}s{stride}_{c_in}->{co}@{oh}x{ow}" + ("+pool" if pool2 else ""), flops, nbytes):
        check(_lib.load().din_conv2d_nhwc_f16(C.byref(d), xp, _p(w_packed), _p(bias), rp, yp, _stream()),
              "din_conv2d_nhwc_f16")
    return out


def conv2d_branches_nhwc(x, w_packed, bias, out, out2, *, split_col, norelu=(0, 0), relu=True, c_in=None, x_c_offset=0,
                         y_c_offset=0, y2_c_offset=0):
    """Several 1x1 convolutions over the same input as ONE GEMM (weight rows stacked): output columns [0, split_col) go to
    `out` at channel y_c_offset, the rest to `out2` at y2_c_offset; columns in norelu = [lo, hi) skip the ReLU
    (din_conv2d_branches_nhwc_f16)."""
    _need(x, torch.float16, "x")
    _need(w_packed, torch.float16, "w_packed")
    _need(out, torch.float16, "out")
    _need(out2, torch.float16, "out2")
    _need(bias, torch.float32, "bias")
    n, h, w, cx = x.shape
    assert w_packed.shape[-3] == 1 and w_packed.shape[-2] == 1
    split = 2 if w_packed.dim() == 5 else 1
    co, ci = w_packed.shape[0], w_packed.shape[-1]
    if c_in is None:
        c_in = min(ci, cx - x_c_offset)
    assert (c_in + 63) // 64 * 64 == ci and x_c_offset + c_in <= cx, (c_in, ci, cx, x_c_offset)
    assert out.shape[:3] == (n, h, w) and out2.shape[:3] == (n, h, w)
    assert y_c_offset + split_col <= out.shape[3] and y2_c_offset + co - split_col <= out2.shape[3]
    d = DinConvDesc(n=n, h=h, w=w, c_in=c_in, x_c_stride=cx, c_out=co, y_c_stride=out.shape[3], kh=1, kw=1, stride=1,
                    pad_h=0, pad_w=0, relu=int(relu), out_f32=0, pool2=0, w_split=split)
    br = DinConvBranchOut(split_col=split_col, y2_c_stride=out2.shape[3], norelu_lo=norelu[0], norelu_hi=norelu[1])
    flops = 2 * n * h * w * co * c_in
    nbytes = 2 * n * h * w * (c_in + co) + 2 * co * ci
    with _launch(f"conv1x1s1_{c_in}->{co}@{h}x{w}+branches", flops, nbytes):
        check(_lib.load().din_conv2d_branches_nhwc_f16(
            C.byref(d), C.byref(br), C.c_void_p(x.data_ptr() + 2 * x_c_offset), _p(w_packed), _p(bias),
            C.c_void_p(out.data_ptr() + 2 * y_c_offset), C.c_void_p(out2.data_ptr() + 2 * y2_c_offset), _stream()),
            "din_conv2d_branches_nhwc_f16")
    return out, out2


def stem_conv(x, w_oihw, bias, *, stride=1, pad=0, relu=True, prep=True):
    """Raw images (0..255) -> prep_images -> conv(+bias,+ReLU) -> NHWC fp16.
    x: fp32 NCHW [n,3,h,w] (the reference loader's tensor) or uint8 NHWC [n,h,w,3] (the decoded frame)."""
    _need(w_oihw, torch.float32, "w_oihw")
    u8 = isinstance(x, torch.Tensor) and x.dtype == torch.uint8
    _need(x, torch.uint8 if u8 else torch.float32, "x")
    if u8:
        n, h, w, c = x.shape
    else:
        n, c, h, w = x.shape
    assert c == 3
    co, ci, kh, kw = w_oihw.shape
    assert ci == 3
    oh = (h + 2 * pad - kh) // stride + 1
    ow = (w + 2 * pad - kw) // stride + 1
    y = torch.empty((n, oh, ow, co), dtype=torch.float16, device=x.device)
    if bias is not None:
        _need(bias, torch.float32, "bias")
    fn = "din_stem_conv_nhwc_u8" if u8 else "din_stem_conv_nchw_f32"
    with _launch(f"stem{kh}x{kw}s{stride}_3->{co}@{oh}x{ow}" + ("_u8" if u8 else ""),
                 2 * n * oh * ow * co * kh * kw * 3, (1 if u8 else 4) * n * 3 * h * w + 2 * n * oh * ow * co):
        check(getattr(_lib.load(), fn)(_p(x), _p(w_oihw), _p(bias), _p(y), n, h, w, co, kh, kw, stride, pad,
                                       int(relu), int(prep), _stream()), fn)
    return y


def stem_pool_supported(x):
    """The fused ResNet-18 stem + max-pool kernel loads image patches by TMA: rows must be 16-byte multiples."""
    if x.data_ptr() % 16 != 0:
        return False
    if x.dtype == torch.uint8:
        return x.dim() == 4 and x.shape[-1] == 3 and x.shape[2] % 16 == 0 and x.shape[1] >= 7
    return x.dim() == 4 and x.shape[1] == 3 and x.shape[3] % 4 == 0 and x.shape[2] >= 7


def stem_conv_pool(x, w_oihw, bias, *, prep=True):
    """Raw images -> prep_images -> conv 7x7 stride 2 pad 3 (+bias) -> ReLU -> MaxPool2d(3, 2, 1) -> NHWC fp16 [n, ph, pw, 64],
    one launch (din_stem7x7_pool_nhwc_f16): resnet18's conv1 / bn1 (folded) / relu / maxpool."""
    _need(w_oihw, torch.float32, "w_oihw")
    u8 = x.dtype == torch.uint8
    _need(x, torch.uint8 if u8 else torch.float32, "x")
    n, h, w = (x.shape[0], x.shape[1], x.shape[2]) if u8 else (x.shape[0], x.shape[2], x.shape[3])
    if tuple(w_oihw.shape) != (64, 3, 7, 7):
        raise _lib.DinError(f"stem_conv_pool: weight must be [64,3,7,7], got {tuple(w_oihw.shape)}")
    oh, ow = (h + 6 - 7) // 2 + 1, (w + 6 - 7) // 2 + 1
    ph, pw = (oh - 1) // 2 + 1, (ow - 1) // 2 + 1
    y = torch.empty((n, ph, pw, 64), dtype=torch.float16, device=x.device)
    if bias is not None:
        _need(bias, torch.float32, "bias")
    with _launch(f"stem7x7s2+pool_3->64@{oh}x{ow}" + ("_u8" if u8 else ""), 2 * n * oh * ow * 64 * 147,
                 (1 if u8 else 4) * n * 3 * h * w + 2 * n * ph * pw * 64):
        check(_lib.load().din_stem7x7_pool_nhwc_f16(_p(x), int(u8), _p(w_oihw), _p(bias), _p(y), n, h, w, int(prep), _stream()),
              "din_stem7x7_pool_nhwc_f16")
    return y


def stem_pair_supported(x):
    """The fused conv1_1 + conv1_2 kernel loads image patches by TMA: rows must be 16-byte multiples."""
    if x.dtype == torch.uint8:
        return x.dim() == 4 and x.shape[-1] == 3 and x.shape[2] % 16 == 0
    return x.dim() == 4 and x.shape[1] == 3 and x.shape[3] % 4 == 0


def stem_conv_pair(x, w1_oihw, b1, w2_packed, b2, *, relu=True, pool2=False, prep=True, out=None):
    """Raw images -> prep_images -> conv3x3(3->64)+bias+ReLU -> conv3x3(64->64)+bias(+ReLU)(+2x2 max-pool) -> NHWC fp16,
    one launch; the 64-channel intermediate never leaves the SM (din_conv3x3_stem_pair_nhwc_f16)."""
    is_u8 = x.dtype == torch.uint8
    if not is_u8:
        _need(x, torch.float32, "x")
    elif not (x.is_cuda and x.is_contiguous()):
        raise _lib.DinError("x must be a contiguous CUDA tensor")
    _need(w1_oihw, torch.float32, "w1_oihw")
    _need(w2_packed, torch.float16, "w2_packed")
    if tuple(w1_oihw.shape) != (64, 3, 3, 3) or tuple(w2_packed.shape) != (64, 3, 3, 64):
        raise _lib.DinError(f"stem_conv_pair: weights must be [64,3,3,3] and packed [64,3,3,64], got "
                            f"{tuple(w1_oihw.shape)} / {tuple(w2_packed.shape)}")
    n, h, w = (x.shape[0], x.shape[1], x.shape[2]) if is_u8 else (x.shape[0], x.shape[2], x.shape[3])
    yh, yw = (h // 2, w // 2) if pool2 else (h, w)
    if out is None:
        out = torch.empty((n, yh, yw, 64), dtype=torch.float16, device=x.device)
    _need(out, torch.float16, "out")
    assert out.shape[:3] == (n, yh, yw) and out.shape[3] >= 64, (tuple(out.shape), (n, yh, yw))
    flops = 2 * n * h * w * 64 * (27 + 9 * 64)
    nbytes = x.numel() * x.element_size() + 2 * n * yh * yw * 64
    with _launch(f"conv3x3s1_3->64->64@{h}x{w}+stem" + ("+pool" if pool2 else ""), flops, nbytes):
        check(_lib.load().din_conv3x3_stem_pair_nhwc_f16(_p(x), int(is_u8), _p(w1_oihw), _p(b1), _p(w2_packed), _p(b2),
                                                         _p(out), n, h, w, out.shape[3], int(relu), int(pool2),
                                                         int(prep), _stream()), "din_conv3x3_stem_pair_nhwc_f16")
    return out


def _pool(fn_name, x, k, stride, pad, out, c, x_c_offset, y_c_offset):
    _need(x, torch.float16, "x")
    n, h, w, cx = x.shape
    c = cx - x_c_offset if c is None else c
    oh = (h + 2 * pad - k) // stride + 1
    ow = (w + 2 * pad - k) // stride + 1
    y = torch.empty((n, oh, ow, c), dtype=torch.float16, device=x.device) if out is None else out
    _need(y, torch.float16, "out")
    assert tuple(y.shape[:3]) == (n, oh, ow) and y_c_offset + c <= y.shape[3] and x_c_offset + c <= cx
    xp = C.c_void_p(x.data_ptr() + 2 * x_c_offset)
    yp = C.c_void_p(y.data_ptr() + 2 * y_c_offset)
    with _launch(f"{fn_name[4:11]}{k}s{stride}_{c}@{oh}x{ow}", 0, 2 * n * c * (h * w + oh * ow)):
        check(getattr(_lib.load(), fn_name)(xp, yp, n, h, w, c, cx, y.shape[3], k, stride, pad, _stream()), fn_name)
    return y


def maxpool2d_nhwc(x, k, stride, pad=0, out=None, c=None, x_c_offset=0, y_c_offset=0):
    return _pool("din_maxpool2d_nhwc_f16", x, k, stride, pad, out, c, x_c_offset, y_c_offset)


def avgpool2d_nhwc(x, k, stride, pad=0, out=None, c=None, x_c_offset=0, y_c_offset=0):
    """count_include_pad=True (torch default)."""
    return _pool("din_avgpool2d_nhwc_f16", x, k, stride, pad, out, c, x_c_offset, y_c_offset)


def avgpool3_bias_relu_nhwc(x, bias, out, *, c, x_c_offset=0, y_c_offset=0, relu=True):
    """out[..., y_c_offset : +c] = relu(avg_pool3x3 s1 p1 (x[..., x_c_offset : +c]) + bias): the tail of a pool branch whose
    1x1 convolution ran first (conv2d_branches_nhwc)."""
    _need(x, torch.float16, "x")
    _need(out, torch.float16, "out")
    _need(bias, torch.float32, "bias")
    n, h, w, cx = x.shape
    assert tuple(out.shape[:3]) == (n, h, w) and y_c_offset + c <= out.shape[3] and x_c_offset + c <= cx
    assert bias.numel() == c
    with _launch(f"avgpool3s1+bias_{c}@{h}x{w}", 0, 4 * n * c * h * w):
        check(_lib.load().din_avgpool3_bias_relu_nhwc_f16(
            C.c_void_p(x.data_ptr() + 2 * x_c_offset), C.c_void_p(out.data_ptr() + 2 * y_c_offset), _p(bias), n, h, w, c,
            cx, out.shape[3], int(relu), _stream()), "din_avgpool3_bias_relu_nhwc_f16")
    return out


def upsample_bilinear_nhwc(x, oh, ow, out=None, c=None, x_c_offset=0, y_c_offset=0):
    """align_corners=True bilinear resize of channels [x_c_offset, x_c_offset+c) into out[..., y_c_offset:]."""
    _need(x, torch.float16, "x")
    n, h, w, cx = x.shape
    c = cx - x_c_offset if c is None else c
    y = torch.empty((n, oh, ow, c), dtype=torch.float16, device=x.device) if out is None else out
    _need(y, torch.float16, "out")
    assert tuple(y.shape[:3]) == (n, oh, ow) and y_c_offset + c <= y.shape[3]
    xp = C.c_void_p(x.data_ptr() + 2 * x_c_offset)
    yp = C.c_void_p(y.data_ptr() + 2 * y_c_offset)
    with _launch(f"upsample_{c}@{oh}x{ow}", 0, 2 * n * c * (h * w + oh * ow)):
        check(_lib.load().din_upsample_bilinear_nhwc_f16(xp, yp, n, h, w, c, cx, y.shape[3], oh, ow, _stream()),
              "din_upsample_bilinear_nhwc_f16")
    return y


# ---------------------------------------------------------------------------------------------
# person-level head
# ---------------------------------------------------------------------------------------------
def roi_align_nhwc(fm, boxes, box_ind, crop_h, crop_w, d=None, out=None, out_f32=False):
    """fm [n_img,h,w,Cs] fp16 NHWC; boxes [m,4] fp32; box_ind [m] int32 -> [m, crop_h*crop_w, d] fp16
    (out_f32: the un-rounded fp32 crops, for the fp32 embedding of very small batches)."""
    _need(fm, torch.float16, "fm")
    _need(boxes, torch.float32, "boxes")
    _need(box_ind, torch.int32, "box_ind")
    n_img, h, w, cs = fm.shape
    d = cs if d is None else d
    m = boxes.shape[0]
    dt = torch.float32 if out_f32 else torch.float16
    if out is None:
        out = torch.empty((m, crop_h * crop_w, d), dtype=dt, device=fm.device)
    _need(out, dt, "out")
    fn = "din_roi_align_nhwc_f16_f32out" if out_f32 else "din_roi_align_nhwc_f16"
    with _launch("roi_align", 0, 2 * n_img * h * w * d + out.element_size() * m * crop_h * crop_w * d):
        check(getattr(_lib.load(), fn)(_p(fm), _p(boxes), _p(box_ind), _p(out), n_img, h, w, d, cs, m,
                                       crop_h, crop_w, _stream()), fn)
    return out


def tmap_cache_stats():
    """(cached maps, hits, misses) of the library's CUtensorMap cache."""
    h, m = C.c_ulonglong(0), C.c_ulonglong(0)
    n = _lib.load().din_tmap_cache_stats(C.byref(h), C.byref(m))
    return n, h.value, m.value


def group_layernorm(x, gamma, beta, *, n_outer, n_inner=1, outer_stride, inner_stride=0, rows=1, row_stride=0,
                    cols, pre=None, post=None, relu=False, eps=1e-5, n_valid=None, out=None):
    _need(x, torch.float32, "x")
    _need(gamma, torch.float32, "gamma")
    _need(beta, torch.float32, "beta")
    if out is None:
        out = torch.empty_like(x)
    with _launch("group_layernorm", 0, 8 * n_outer * n_inner * rows * cols):
        check(_lib.load().din_group_layernorm_f32(_p(x), _p(pre), _p(post), _p(gamma), _p(beta), _p(out), n_outer,
                                                  n_inner, outer_stride, inner_stride, rows, row_stride, cols,
                                                  float(eps), int(relu), _p(n_valid), _stream()),
              "din_group_layernorm_f32")
    return out


def linear_f32(x, w, bias=None, *, relu=False, out=None, accumulate=False):
    """x [..., k] fp32, w [n, k] fp32 -> [..., n]."""
    _need(x, torch.float32, "x")
    _need(w, torch.float32, "w")
    k = x.shape[-1]
    n = w.shape[0]
    m = x.numel() // k
    if out is None:
        assert not accumulate
        out = torch.empty(x.shape[:-1] + (n,), dtype=torch.float32, device=x.device)
    with _launch(f"linear_f32_{k}->{n}", 2 * m * n * k, 4 * (m * k + n * k + m * n)):
        check(_lib.load().din_linear_f32(_p(x), _p(w), _p(bias), _p(out), m, n, k, int(relu), int(accumulate),
                                         _stream()), "din_linear_f32")
    return out


def pack_din_weights(p_w, p_b, s_w=None, s_b=None):
    """OIHW p_conv [2k2,C,kt,kn] (+ scale_conv [k2,C,kt,kn]) -> tap-major [kt*kn][n_out][C], bias [n_out].
    Pure layout change (done once per weight update, not on the hot path)."""
    ws = [p_w] + ([s_w] if s_w is not None else [])
    bs = [p_b] + ([s_b] if s_b is not None else [])
    w = torch.cat(ws, dim=0)                                   # [n_out, C, kt, kn]
    n_out, c, kt, kn = w.shape
    w_tap = w.permute(2, 3, 0, 1).reshape(kt * kn, n_out, c).contiguous().float()
    return w_tap, torch.cat(bs, dim=0).contiguous().float()


def dynamic_infer(x, w_tap, b_cat, kernel, ratio, *, scale_factor=True, out=None, coef=1.0, coef_ptr=None,
                  accumulate=False, n_valid=None):
    """x [b,t,n,c] fp32 -> y [b,t,n,c]; y = coef*DIN_ratio(x) (or += when accumulate)."""
    _need(x, torch.float32, "x")
    _need(w_tap, torch.float32, "w_tap")
    _need(b_cat, torch.float32, "b_cat")
    b, t, n, c = x.shape
    kt, kn = kernel
    if out is None:
        assert not accumulate
        # actors beyond n_valid are never written: start from zeros so they are well defined
        out = torch.zeros_like(x) if n_valid is not None else torch.empty_like(x)
    with _launch(f"dynamic_infer_k{kt}x{kn}_r{ratio}_c{c}", 2 * b * t * n * w_tap.shape[1] * kt * kn * c,
                 8 * b * t * n * c + 4 * w_tap.numel()):
        check(_lib.load().din_dynamic_infer_f32(_p(x), _p(w_tap), _p(b_cat), _p(out), b, t, n, c, kt, kn, ratio,
                                                int(scale_factor), C.c_void_p(coef_ptr or 0), float(coef),
                                                int(accumulate), _p(n_valid), _stream()), "din_dynamic_infer_f32")
    return out


def readout(s, w, bias, n_valid=None):
    """s [b,t,n,c] fp32 -> logits [b,a] = mean_t fc(max_n s)."""
    _need(s, torch.float32, "s")
    _need(w, torch.float32, "w")
    _need(bias, torch.float32, "bias")
    b, t, n, c = s.shape
    a = w.shape[0]
    out = torch.empty((b, a), dtype=torch.float32, device=s.device)
    with _launch("readout", 0, 4 * b * t * n * c):
        check(_lib.load().din_readout_f32(_p(s), _p(w), _p(bias), _p(out), b, t, n, c, a, _p(n_valid), _stream()),
              "din_readout_f32")
    return out


# ---------------------------------------------------------------------------------------------
# after the path: loss / metrics on the device, stage-1 helper
# ---------------------------------------------------------------------------------------------
def ce_metrics(logits, labels, *, class_weight=None, loss_scale=1.0, conf=None, meters=None, want_grad=False):
    """One launch: (loss [1], correct [1] int32, dlogits or None).  conf [a,a] int32 and meters [4] fp64 accumulate."""
    _need(logits, torch.float32, "logits")
    _need(labels, torch.int64, "labels")
    b, a = logits.shape
    assert labels.shape == (b,), (labels.shape, b)
    if class_weight is not None:
        _need(class_weight, torch.float32, "class_weight")
        assert class_weight.shape == (a,)
    if conf is not None:
        _need(conf, torch.int32, "conf")
        assert conf.shape == (a, a)
    if meters is not None:
        _need(meters, torch.float64, "meters")
        assert meters.shape == (4,)
    loss = torch.empty((1,), dtype=torch.float32, device=logits.device)
    correct = torch.empty((1,), dtype=torch.int32, device=logits.device)
    dlogits = torch.empty_like(logits) if want_grad else None
    with _launch("ce_metrics", 0, 4 * b * a * (2 if want_grad else 1)):
        check(_lib.load().din_ce_metrics_f32(_p(logits), _p(labels), _p(class_weight), float(loss_scale), _p(loss),
                                             _p(correct), _p(conf), _p(meters), _p(dlogits), b, a, _stream()),
              "din_ce_metrics_f32")
    return loss, correct, dlogits


def mean_axis(x, dim):
    """fp32 mean over one axis (kept dims removed)."""
    _need(x, torch.float32, "x")
    shape = list(x.shape)
    outer = 1
    for s in shape[:dim]:
        outer *= s
    inner = 1
    for s in shape[dim + 1:]:
        inner *= s
    y = torch.empty(shape[:dim] + shape[dim + 1:], dtype=torch.float32, device=x.device)
    with _launch("mean_axis", 0, 4 * (x.numel() + y.numel())):
        check(_lib.load().din_mean_axis_f32(_p(x), _p(y), outer, shape[dim], inner, _stream()), "din_mean_axis_f32")
    return y


# ---------------------------------------------------------------------------------------------
# backward of the person-level head (first slice)
# ---------------------------------------------------------------------------------------------
def gemm_f32(a, b, *, m, n, k, a_strides, b_strides, out=None, alpha=1.0, accumulate=False):
    """out[m,n] (+)= alpha * sum_k A(m,k) B(k,n) with explicit element strides (a fp32; b fp32 or fp16)."""
    _need(a, torch.float32, "a")
    if b.dtype not in (torch.float32, torch.float16) or not b.is_cuda or not b.is_contiguous():
        raise _lib.DinError("gemm_f32: b must be a contiguous CUDA fp32/fp16 tensor")
    if out is None:
        assert not accumulate
        out = torch.empty((m, n), dtype=torch.float32, device=a.device)
    _need(out, torch.float32, "out")
    with _launch(f"gemm_f32_{m}x{n}x{k}", 2 * m * n * k, 4 * (m * k + m * n) + b.element_size() * n * k):
        check(_lib.load().din_gemm_f32(_p(a), a_strides[0], a_strides[1], _p(b), int(b.dtype == torch.float16),
                                       b_strides[0], b_strides[1], _p(out), out.stride(0) if out.dim() == 2 else n,
                                       m, n, k, float(alpha), int(accumulate), _stream()), "din_gemm_f32")
    return out


def linear_bwd(x, w, dy, *, need_dx=True, dx_out=None, dx_accumulate=False, has_bias=True):
    """y = x.w^T (+ bias), x [m,k] (fp32, or fp16 crops), w [n,k] (None when need_dx is False), dy [m,n]
    -> (dx [m,k] | None, dw [n,k], db [n] | None)."""
    n = dy.shape[-1]
    m = dy.numel() // n
    k = x.numel() // m
    assert w is None or tuple(w.shape) == (n, k), (None if w is None else tuple(w.shape), n, k)
    dx = None
    if need_dx:
        # dx[m,k] = sum_n dy[m,n] w[n,k]
        dx = gemm_f32(dy, w, m=m, n=k, k=n, a_strides=(n, 1), b_strides=(k, 1), out=dx_out, accumulate=dx_accumulate)
    # dw[n,k] = sum_m dy[m,n] x[m,k]
    dw = gemm_f32(dy, x, m=n, n=k, k=m, a_strides=(1, n), b_strides=(k, 1))
    db = None
    if has_bias:
        db = torch.empty((n,), dtype=torch.float32, device=dy.device)
        with _launch("colsum", 0, 4 * m * n):
            check(_lib.load().din_colsum_f32(_p(dy), _p(db), m, n, n, _stream()), "din_colsum_f32")
    return dx, dw, db


def scale_mask(x, mask, scale, out=None):
    """x * mask * scale (dropout forward / backward); mask uint8 0/1 or None."""
    _need(x, torch.float32, "x")
    if mask is not None:
        _need(mask, torch.uint8, "mask")
        assert mask.numel() == x.numel()
    if out is None:
        out = torch.empty_like(x)
    with _launch("scale_mask", 0, 9 * x.numel()):
        check(_lib.load().din_scale_mask_f32(_p(x), _p(mask), float(scale), _p(out), x.numel(), _stream()),
              "din_scale_mask_f32")
    return out


def relu_bwd_f32(y, dy):
    """dy * [y > 0] (fp32)."""
    _need(y, torch.float32, "y")
    _need(dy, torch.float32, "dy")
    assert y.numel() == dy.numel()
    dz = torch.empty_like(dy)
    with _launch("relu_bwd_f32", 0, 12 * y.numel()):
        check(_lib.load().din_relu_bwd_f32(_p(y), _p(dy), _p(dz), y.numel(), _stream()), "din_relu_bwd_f32")
    return dz


def readout_bwd(s, w, dlogits, n_valid=None):
    """-> (ds [b,t,n,c], dw [a,c], dbias [a])."""
    _need(s, torch.float32, "s")
    _need(w, torch.float32, "w")
    _need(dlogits, torch.float32, "dlogits")
    b, t, n, c = s.shape
    a = w.shape[0]
    ds = torch.empty_like(s)
    ws = torch.empty((b, t, c), dtype=torch.float32, device=s.device)
    dw = torch.empty_like(w)
    db = torch.empty((a,), dtype=torch.float32, device=s.device)
    global LAUNCHES
    LAUNCHES += 1                                   # two kernels
    with _launch("readout_bwd", 0, 8 * s.numel()):
        check(_lib.load().din_readout_bwd_f32(_p(s), _p(w), _p(dlogits), _p(ds), _p(ws), _p(dw), _p(db), b, t, n, c, a,
                                              _p(n_valid), _stream()), "din_readout_bwd_f32")
    return ds, dw, db


def group_layernorm_bwd(x, gamma, beta, dy, *, n_outer, n_inner=1, outer_stride, inner_stride=0, rows=1,
                        row_stride=0, cols, pre=None, relu=False, eps=1e-5, n_valid=None, dx_out=None,
                        accumulate=False, param_grads=True):
    """-> (dx, dgamma, dbeta); dx is also the gradient of `pre`; the gradient of `post` is dy itself."""
    _need(x, torch.float32, "x")
    _need(dy, torch.float32, "dy")
    if dx_out is None:
        assert not accumulate
        # groups skipped through n_valid are never written: start from zeros so they are well defined
        dx_out = torch.zeros_like(x) if n_valid is not None else torch.empty_like(x)
    stats = torch.empty((n_outer * n_inner, 2), dtype=torch.float32, device=x.device)
    dg = torch.empty_like(gamma) if param_grads else None
    db = torch.empty_like(beta) if param_grads else None
    global LAUNCHES
    LAUNCHES += 1 if param_grads else 0             # second kernel: parameter gradients
    with _launch("group_layernorm_bwd", 0, 16 * n_outer * n_inner * rows * cols):
        check(_lib.load().din_group_layernorm_bwd_f32(_p(x), _p(pre), _p(gamma), _p(beta), _p(dy), _p(dx_out), _p(dg),
                                                      _p(db), _p(stats), n_outer, n_inner, outer_stride, inner_stride,
                                                      rows, row_stride, cols, float(eps), int(relu), int(accumulate),
                                                      _p(n_valid), _stream()), "din_group_layernorm_bwd_f32")
    return dx_out, dg, db


def dynamic_infer_bwd(x, w_tap, b_cat, dy, dx, kernel, ratio, *, scale_factor=True, coef=1.0, coef_ptr=None,
                      want_dcoef=False, n_valid=None):
    """dx += ...; returns (dw_tap, db_cat, dcoef | None).  Launches four kernels."""
    _need(x, torch.float32, "x")
    _need(dy, torch.float32, "dy")
    _need(dx, torch.float32, "dx")
    b, t, n, c = x.shape
    kt, kn = kernel
    n_out = w_tap.shape[1]
    dw = torch.empty_like(w_tap)
    db = torch.empty_like(b_cat)
    dcoef = torch.empty((1,), dtype=torch.float32, device=x.device) if want_dcoef else None
    ws_floats = _lib.load().din_dynamic_infer_bwd_ws_floats(b, t, n, c, kt, kn, int(scale_factor))
    ws = torch.empty((ws_floats,), dtype=torch.float32, device=x.device)
    global LAUNCHES
    LAUNCHES += 3                                   # this entry point launches four kernels
    with _launch(f"dynamic_infer_bwd_k{kt}x{kn}_r{ratio}_c{c}", 6 * b * t * n * n_out * kt * kn * c, 16 * x.numel()):
        check(_lib.load().din_dynamic_infer_bwd_f32(_p(x), _p(w_tap), _p(b_cat), _p(dy), _p(dx), _p(dw), _p(db),
                                                    _p(dcoef), _p(ws), b, t, n, c, kt, kn, ratio, int(scale_factor),
                                                    C.c_void_p(coef_ptr or 0), float(coef), _p(n_valid), _stream()),
              "din_dynamic_infer_bwd_f32")
    return dw, db, dcoef


# ---------------------------------------------------------------------------------------------
# backward of the backbone (VGG-16: 3x3 stride-1 convolutions)
# ---------------------------------------------------------------------------------------------
def conv2d_wgrad_nhwc(x, dz, dw, dbias=None, *, pad=(1, 1), inv_scale=None, c_in=None, c_out=None, x_c_offset=0,
                      dz_c_offset=0):
    """dw [c_out,kh,kw,c_in] fp32 (+)= sum_pixels dz (x) x_shifted;  dbias [c_out] (+)= sum_pixels dz.  Accumulates.
    x / dz may be channel slices (x_c_offset / dz_c_offset) of wider NHWC buffers."""
    _need(x, torch.float16, "x")
    _need(dz, torch.float16, "dz")
    _need(dw, torch.float32, "dw")
    n, h, w, cx = x.shape
    co, kh, kw, ci = dw.shape
    assert dz.shape[0] == n and dz.shape[1] == h + 2 * pad[0] - kh + 1 and dz.shape[2] == w + 2 * pad[1] - kw + 1, \
        (tuple(x.shape), tuple(dz.shape))
    assert x_c_offset + ci <= cx and dz_c_offset + co <= dz.shape[3]
    if dbias is not None:
        _need(dbias, torch.float32, "dbias")
    if inv_scale is not None:
        _need(inv_scale, torch.float32, "inv_scale")
    flops = 2 * n * dz.shape[1] * dz.shape[2] * co * kh * kw * ci
    with _launch(f"wgrad{kh}x{kw}_{ci}->{co}@{dz.shape[1]}x{dz.shape[2]}", flops, 2 * (x.numel() + dz.numel())):
        check(_lib.load().din_conv2d_wgrad_nhwc_f16(C.c_void_p(x.data_ptr() + 2 * x_c_offset),
                                                    C.c_void_p(dz.data_ptr() + 2 * dz_c_offset), _p(dw), _p(dbias),
                                                    _p(inv_scale), n, h, w, ci, cx, co,
                                                    dz.shape[3], kh, kw, pad[0], pad[1], _stream()),
              "din_conv2d_wgrad_nhwc_f16")
    return dw


def roi_align_bwd(dcrops, boxes, box_ind, dfm, crop_h, crop_w, d=None):
    """dfm [n_img,h,w,Cs] fp32 += scatter(dcrops [m, crop_h*crop_w, d])."""
    _need(dcrops, torch.float32, "dcrops")
    _need(dfm, torch.float32, "dfm")
    _need(boxes, torch.float32, "boxes")
    _need(box_ind, torch.int32, "box_ind")
    n_img, h, w, cs = dfm.shape
    d = cs if d is None else d
    m = boxes.shape[0]
    assert dcrops.numel() == m * crop_h * crop_w * d
    with _launch("roi_align_bwd", 0, 4 * (dcrops.numel() * 5)):
        check(_lib.load().din_roi_align_bwd_f32(_p(dcrops), _p(boxes), _p(box_ind), _p(dfm), n_img, h, w, d, cs, m,
                                                crop_h, crop_w, _stream()), "din_roi_align_bwd_f32")
    return dfm


def grad_to_f16(x, target=256.0):
    """-> (x * S as fp16, scale_ws [4] fp32 device: [1] = S, [2] = 1/S).  Two kernels."""
    _need(x, torch.float32, "x")
    y = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    ws = torch.empty((4,), dtype=torch.float32, device=x.device)
    global LAUNCHES
    LAUNCHES += 1
    with _launch("grad_to_f16", 0, 10 * x.numel()):
        check(_lib.load().din_grad_to_f16(_p(x), _p(y), _p(ws), x.numel(), float(target), _stream()), "din_grad_to_f16")
    return y, ws


def relu_pool_bwd_nhwc(y, dy, pool):
    """y [n,h,w,c] fp16 saved ReLU output; dy [n,h,w,c] (or [n,h/2,w/2,c] with pool) -> dz [n,h,w,c]."""
    _need(y, torch.float16, "y")
    _need(dy, torch.float16, "dy")
    n, h, w, c = y.shape
    assert tuple(dy.shape) == ((n, h // 2, w // 2, c) if pool else (n, h, w, c)), (tuple(y.shape), tuple(dy.shape), pool)
    dz = torch.empty_like(y)
    with _launch("relu_pool_bwd" if pool else "relu_bwd", 0, 2 * (2 * y.numel() + dy.numel())):
        check(_lib.load().din_relu_pool_bwd_nhwc_f16(_p(y), _p(dy), _p(dz), n, h, w, c, int(pool), _stream()),
              "din_relu_pool_bwd_nhwc_f16")
    return dz


def stem_wgrad(x, dz, dw, dbias, *, stride=1, pad=1, inv_scale=None, prep=True):
    """Stem weight gradient: dw [64,3,k,k] fp32 (OIHW), dbias [64] (+)= ...; x raw frames (fp32 NCHW or uint8 NHWC).
    VGG-16: k = 3, stride 1, pad 1;  ResNet-18: k = 7, stride 2, pad 3."""
    u8 = x.dtype == torch.uint8
    _need(x, torch.uint8 if u8 else torch.float32, "x")
    _need(dz, torch.float16, "dz")
    _need(dw, torch.float32, "dw")
    n, h, w = (x.shape[0], x.shape[1], x.shape[2]) if u8 else (x.shape[0], x.shape[2], x.shape[3])
    k, co = dw.shape[2], dw.shape[0]
    oh, ow = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    assert tuple(dz.shape) == (n, oh, ow, co) and tuple(dw.shape) == (co, 3, k, k), (tuple(dz.shape), tuple(dw.shape))
    with _launch(f"stem_wgrad{k}x{k}", 2 * n * oh * ow * co * 3 * k * k, x.numel() * x.element_size() + 2 * dz.numel()):
        check(_lib.load().din_stem_wgrad(_p(x), int(u8), _p(dz), _p(dw), _p(dbias), _p(inv_scale), n, h, w, co, k, k,
                                         stride, pad, int(prep), _stream()), "din_stem_wgrad")
    return dw


def scatter2_nhwc(src, h, w, dst=None):
    """dst[n, 2oy, 2ox] (+)= src[n, oy, ox]: dst None -> zero insertion into a new [n,h,w,c]; else accumulate."""
    _need(src, torch.float16, "src")
    n, oh, ow, c = src.shape
    acc = dst is not None
    if dst is None:
        dst = torch.empty((n, h, w, c), dtype=torch.float16, device=src.device)
    _need(dst, torch.float16, "dst")
    assert tuple(dst.shape) == (n, h, w, c)
    with _launch("scatter2", 0, 2 * (src.numel() + dst.numel())):
        check(_lib.load().din_scatter2_nhwc_f16(_p(src), _p(dst), n, h, w, c, oh, ow, int(acc), _stream()),
              "din_scatter2_nhwc_f16")
    return dst


def add_f16(a, b):
    _need(a, torch.float16, "a")
    _need(b, torch.float16, "b")
    assert a.shape == b.shape
    y = torch.empty_like(a)
    with _launch("add_f16", 0, 6 * a.numel()):
        check(_lib.load().din_add_f16(_p(a), _p(b), _p(y), a.numel(), _stream()), "din_add_f16")
    return y


def maxpool3s2_relu_bwd_nhwc(x, dy):
    """x [n,h,w,c] saved ReLU output, dy [n,oh,ow,c] -> dz [n,h,w,c] (MaxPool2d(3,2,1) + ReLU backward)."""
    _need(x, torch.float16, "x")
    _need(dy, torch.float16, "dy")
    n, h, w, c = x.shape
    assert tuple(dy.shape) == (n, (h - 1) // 2 + 1, (w - 1) // 2 + 1, c), (tuple(x.shape), tuple(dy.shape))
    dz = torch.empty_like(x)
    with _launch("maxpool3s2_bwd", 0, 2 * (2 * x.numel() + dy.numel())):
        check(_lib.load().din_maxpool3s2_relu_bwd_nhwc_f16(_p(x), _p(dy), _p(dz), n, h, w, c, _stream()),
              "din_maxpool3s2_relu_bwd_nhwc_f16")
    return dz


def maxpool3s2_bwd_nhwc(x, dy, dz, *, c, pad=0, x_c_offset=0, dy_c_offset=0, dz_c_offset=0, accumulate=False):
    """dz[..., dz_c_offset:+c] (+)= MaxPool2d(3, 2, pad) backward of dy[..., dy_c_offset:+c] routed by x[..., x_c_offset:+c]
    (first maximum of each window), times [x > 0]."""
    for name, t in (("x", x), ("dy", dy), ("dz", dz)):
        _need(t, torch.float16, name)
    n, h, w, cx = x.shape
    oh, ow = (h + 2 * pad - 3) // 2 + 1, (w + 2 * pad - 3) // 2 + 1
    assert tuple(dy.shape[:3]) == (n, oh, ow) and tuple(dz.shape[:3]) == (n, h, w)
    assert x_c_offset + c <= cx and dy_c_offset + c <= dy.shape[3] and dz_c_offset + c <= dz.shape[3]
    with _launch("maxpool3s2_bwd", 0, 2 * c * n * (2 * h * w + oh * ow)):
        check(_lib.load().din_maxpool3s2_bwd_nhwc_f16(
            C.c_void_p(x.data_ptr() + 2 * x_c_offset), C.c_void_p(dy.data_ptr() + 2 * dy_c_offset),
            C.c_void_p(dz.data_ptr() + 2 * dz_c_offset), n, h, w, c, cx, dy.shape[3], dz.shape[3], pad, int(accumulate),
            _stream()), "din_maxpool3s2_bwd_nhwc_f16")
    return dz


def relu_bwd_slice_nhwc(y, dy, dz, *, c, y_c_offset=0, dy_c_offset=0, dz_c_offset=0):
    """dz[..., dz_c_offset:+c] = dy[..., dy_c_offset:+c] * [y[..., y_c_offset:+c] > 0]  (fp16 NHWC, same pixels)."""
    for name, t in (("y", y), ("dy", dy), ("dz", dz)):
        _need(t, torch.float16, name)
    rows = y.numel() // y.shape[-1]
    assert dy.numel() // dy.shape[-1] == rows == dz.numel() // dz.shape[-1]
    assert y_c_offset + c <= y.shape[-1] and dy_c_offset + c <= dy.shape[-1] and dz_c_offset + c <= dz.shape[-1]
    with _launch("relu_bwd_slice", 0, 6 * rows * c):
        check(_lib.load().din_relu_bwd_slice_nhwc_f16(
            C.c_void_p(y.data_ptr() + 2 * y_c_offset), C.c_void_p(dy.data_ptr() + 2 * dy_c_offset),
            C.c_void_p(dz.data_ptr() + 2 * dz_c_offset), rows, c, y.shape[-1], dy.shape[-1], dz.shape[-1], _stream()),
            "din_relu_bwd_slice_nhwc_f16")
    return dz


def upsample_bilinear_bwd_nhwc(dy, h, w, *, c, dy_c_offset=0, out=None):
    """Backward of upsample_bilinear_nhwc: dy [n,oh,ow,*] (channels dy_c_offset:+c) -> dx [n,h,w,c]."""
    _need(dy, torch.float16, "dy")
    n, oh, ow, cy = dy.shape
    if out is None:
        out = torch.empty((n, h, w, c), dtype=torch.float16, device=dy.device)
    _need(out, torch.float16, "out")
    assert tuple(out.shape[:3]) == (n, h, w) and out.shape[3] >= c and dy_c_offset + c <= cy
    with _launch("upsample_bwd", 0, 2 * n * c * (h * w + oh * ow)):
        check(_lib.load().din_upsample_bilinear_bwd_nhwc_f16(C.c_void_p(dy.data_ptr() + 2 * dy_c_offset), _p(out), n, h, w, c,
                                                             cy, out.shape[3], oh, ow, _stream()),
              "din_upsample_bilinear_bwd_nhwc_f16")
    return out


def colsum_nhwc(dz, db, *, c, c_offset=0, inv_scale=None):
    """db[0:c] += inv_scale * sum over pixels of dz[..., c_offset:+c]  (fp16 in, fp32 accumulate)."""
    _need(dz, torch.float16, "dz")
    _need(db, torch.float32, "db")
    rows = dz.numel() // dz.shape[-1]
    assert c_offset + c <= dz.shape[-1] and db.numel() == c
    with _launch("colsum", 0, 2 * rows * c):
        check(_lib.load().din_colsum_nhwc_f16(C.c_void_p(dz.data_ptr() + 2 * c_offset), _p(db), rows, c, dz.shape[-1],
                                              _p(inv_scale), _stream()), "din_colsum_nhwc_f16")
    return db


def bn_gamma_grad(dz, zsrc, gamma, beta, dgamma, *, sub=None, inv_scale=None):
    """dgamma [c] fp32 += inv_scale * sum dz * ((zsrc - sub) - beta) / gamma   (eval-mode BN folded into its conv)."""
    _need(dz, torch.float16, "dz")
    _need(zsrc, torch.float16, "zsrc")
    _need(dgamma, torch.float32, "dgamma")
    c = dz.shape[-1]
    rows = dz.numel() // c
    assert zsrc.shape == dz.shape and (sub is None or sub.shape == dz.shape)
    with _launch("bn_gamma_grad", 0, 2 * dz.numel() * (3 if sub is not None else 2)):
        check(_lib.load().din_bn_gamma_grad_f16(_p(dz), _p(zsrc), _p(sub), _p(gamma), _p(beta), _p(dgamma), _p(inv_scale),
                                                rows, c, _stream()), "din_bn_gamma_grad_f16")
    return dgamma


def bn_fold_grads(w, dwf, dbeta, running_mean, running_var, dgamma, eps=1e-5):
    """dgamma[r] = rsqrt(var[r] + eps) * (<w[r], dwf[r]> - mean[r] * dbeta[r])   (eval-mode BN folded into its conv;
    w the un-folded weight, dwf the gradient w.r.t. the folded one, same element order)."""
    for name, t in (("w", w), ("dwf", dwf), ("dbeta", dbeta), ("running_mean", running_mean),
                    ("running_var", running_var), ("dgamma", dgamma)):
        _need(t, torch.float32, name)
    rows = w.shape[0]
    assert w.shape == dwf.shape and dbeta.numel() == rows == dgamma.numel() == running_mean.numel()
    with _launch("bn_fold_grads", 2 * w.numel(), 8 * w.numel()):
        check(_lib.load().din_bn_fold_grads_f32(_p(w), _p(dwf), _p(dbeta), _p(running_mean), _p(running_var), float(eps),
                                                _p(dgamma), rows, w.numel() // rows, _stream()), "din_bn_fold_grads_f32")
    return dgamma


def scale_rows(w, scale):
    """w[r] *= scale[r] in place (w fp32 [rows, ...])."""
    _need(w, torch.float32, "w")
    _need(scale, torch.float32, "scale")
    rows = w.shape[0]
    with _launch("scale_rows", 0, 8 * w.numel()):
        check(_lib.load().din_scale_rows_f32(_p(w), _p(scale), rows, w.numel() // rows, _stream()), "din_scale_rows_f32")
    return w


# ---- BatchNorm2d on batch statistics (csrc/bn_train.cu) ----------------------------------------------------------------
def bn_train_forward(z, gamma, beta, running_mean=None, running_var=None, *, eps=1e-5, momentum=0.1, residual=None,
                     relu=True, out=None):
    """z: raw convolution output fp16 or fp32 [..., c].  -> (y = [relu](BN_batch(z) [+ residual]) fp16, (mean, invstd)
    fp32 [c]).  running_mean / running_var are updated in place as nn.BatchNorm2d.train() does."""
    f32 = z.dtype == torch.float32
    _need(z, torch.float32 if f32 else torch.float16, "z")
    c = z.shape[-1]
    rows = z.numel() // c
    for name, t in (("gamma", gamma), ("beta", beta), ("running_mean", running_mean), ("running_var", running_var)):
        if t is not None:
            _need(t, torch.float32, name)
    sums = torch.empty((2, c), dtype=torch.float64, device=z.device)
    par = torch.empty((4, c), dtype=torch.float32, device=z.device)          # scale, shift, mean, invstd
    y = torch.empty(z.shape, dtype=torch.float16, device=z.device) if out is None else out
    _need(y, torch.float16, "out")
    assert y.shape == z.shape
    if residual is not None:
        _need(residual, torch.float16, "residual")
        assert residual.shape == z.shape
    lib = _lib.load()
    with _launch("bn_stats", 0, z.element_size() * z.numel()):
        check(lib.din_bn_stats(_p(z), int(f32), rows, c, _p(sums[0]), _p(sums[1]), _stream()), "din_bn_stats")
    check(lib.din_bn_finalize_f32(_p(sums[0]), _p(sums[1]), rows, _p(gamma), _p(beta), float(eps), float(momentum),
                                  _p(running_mean), _p(running_var), _p(par[0]), _p(par[1]), _p(par[2]), _p(par[3]), c,
                                  _stream()), "din_bn_finalize_f32")
    with _launch("bn_apply", 0, z.numel() * (z.element_size() + (4 if residual is not None else 2))):
        check(lib.din_bn_apply(_p(z), int(f32), _p(par[0]), _p(par[1]), _p(residual), _p(y), rows, c, int(relu),
                               _stream()), "din_bn_apply")
    return y, (par[2], par[3])


def bn_train_backward(g, z, mean, invstd, gamma, dbeta=None, dgamma=None, *, inv_scale=None):
    """g: dY fp16 (masked by the ReLU, times the loss scale) -> dz fp16; dbeta / dgamma [c] fp32 += (unscaled) sums."""
    f32 = z.dtype == torch.float32
    _need(g, torch.float16, "g")
    _need(z, torch.float32 if f32 else torch.float16, "z")
    assert g.shape == z.shape
    c = z.shape[-1]
    rows = z.numel() // c
    sums = torch.empty(2 * c, dtype=torch.float32, device=z.device)
    dz = torch.empty_like(g)
    with _launch("bn_bwd", 0, z.numel() * (2 * z.element_size() + 6)):
        check(_lib.load().din_bn_bwd(_p(g), _p(z), int(f32), _p(mean), _p(invstd), _p(gamma), _p(sums), _p(dz), _p(dbeta),
                                     _p(dgamma), _p(inv_scale), rows, c, _stream()), "din_bn_bwd")
    return dz


def pack_flat(tensors, flat, offsets, scale=1.0):
    """flat[offsets[i] : offsets[i] + tensors[i].numel()] = scale * tensors[i] for all i, in ONE launch (csrc/flat.cu)."""
    _need(flat, torch.float32, "flat")
    arr = (_lib.DinFlatJob * len(tensors))()
    total = 0
    for j, t, off in zip(arr, tensors, offsets):
        _need(t, torch.float32, "gradient")
        if off < 0 or off + t.numel() > flat.numel():
            raise _lib.DinError("pack_flat: tensor does not fit the flat buffer")
        j.src, j.dst_offset, j.numel = t.data_ptr(), off, t.numel()
        total += t.numel()
    with _launch(f"pack_flat_{len(tensors)}", 0, 8 * total):
        check(_lib.load().din_pack_flat_f32(arr, len(tensors), _p(flat), float(scale), _stream()), "din_pack_flat_f32")
    return flat


def context_attention(q, img, posbias, n):
    """q [H, M, 128] fp32, img [F, P, H*128] fp32, posbias [P, H*128] fp32 -> ctx [H, M, 128] (M = F * n): per head and
    actor, softmax attention over the frame's P map positions (din_context_attention_f32)."""
    _need(q, torch.float32, "q")
    _need(img, torch.float32, "img")
    _need(posbias, torch.float32, "posbias")
    heads, m, d = q.shape
    frames, pixels = img.shape[0], img.shape[1]
    assert d == 128 and img.shape[2] == heads * 128 and tuple(posbias.shape) == (pixels, heads * 128) and m == frames * n
    ctx = torch.empty_like(q)
    with _launch(f"context_attention_{pixels}px", 4 * m * heads * pixels * 128, 4 * (img.numel() * 2 + 2 * q.numel())):
        check(_lib.load().din_context_attention_f32(_p(q), _p(img), _p(posbias), _p(ctx), frames, n, pixels, heads,
                                                    _stream()), "din_context_attention_f32")
    return ctx


def context_attention_bwd(q, img, posbias, dctx, dq_add=None):
    """Backward of context_attention: -> (dq [H, M, 128] (+ dq_add), dimg [F, P, H*128])."""
    for name, t in (("q", q), ("img", img), ("posbias", posbias), ("dctx", dctx)) + ((("dq_add", dq_add),) if dq_add is not None else ()):
        _need(t, torch.float32, name)
    heads, m, d = q.shape
    frames, pixels = img.shape[0], img.shape[1]
    n = m // frames
    assert d == 128 and dctx.shape == q.shape and img.shape[2] == heads * 128 and m == frames * n
    dq, dimg = torch.empty_like(q), torch.empty_like(img)
    with _launch(f"context_attention_bwd_{pixels}px", 10 * m * heads * pixels * 128, 4 * (img.numel() * 4 + 4 * q.numel())):
        check(_lib.load().din_context_attention_bwd_f32(_p(q), _p(img), _p(posbias), _p(dctx), _p(dq_add), _p(dq), _p(dimg),
                                                        frames, n, pixels, heads, _stream()), "din_context_attention_bwd_f32")
    return dq, dimg
