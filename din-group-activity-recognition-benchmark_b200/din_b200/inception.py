"""inception.py — the truncated Inception-v3 backbone plan (reference backbone/backbone.py:10-85) plus the
multiscale build of infer_model.py:165-172, on NHWC fp16.

torchvision's BasicConv2d = Conv2d(bias=False) + BatchNorm2d(eps=1e-3) + ReLU: BN (eval statistics) is
folded into the packed weights' scale and the epilogue bias.  Every branch convolution writes its output
channels directly at its offset of the block's concat buffer (`y_c_offset` / `y_c_stride` of the conv ABI),
so `torch.cat` never runs; Mixed_5d writes into channels [0, 288) of the multiscale map and Mixed_6e's
768 channels are bilinearly upsampled (align_corners=True) into channels [288, 1056) of the same map.
Channel counts that are not multiples of 64 (32, 48, 80, 96, 160, 288) need no padded buffers: the conv's
TMA tensor map is given the real channel extent and zero-fills the rest of the last 64-channel K block.
"""
from __future__ import annotations

import os

import torch

from . import ops
from .engine import _Conv, _Stem, _fold_bn

D_OUT = 288 + 768            # cfg.emb_features for inv3 (config.py:41)
D_STRIDE = (D_OUT + 63) // 64 * 64   # 1088: channel stride of the multiscale map (pad channels stay zero)


def _c(v, k, s, p=0):
    return (v + 2 * p - k) // s + 1


def _split_for(name):
    """2 = hi + lo fp16 weight parts (exact weights, 2x tensor work), 1 = one fp16 part (error-feedback rounding).
    DIN_INV3_SPLIT (measurement knob): 'none' (default), 'all', or a comma list of substrings of the layer names that
    keep the split.  tests/tools/inv3_split_study.py (profiles/inv3_split_study_r1.log): with error-feedback weight rounding
    in place the logits error is set by the fp16 ACTIVATION rounding of the 37-conv chain -- 5.1e-4 .. 8.8e-4 of
    max|logit| without any split vs 5.9e-4 .. 7.6e-4 with all layers split (4 cases, tolerance 1e-3) -- while the
    split costs 2x tensor work: 7.70 -> 5.48 ms per 16 frames at 720p."""
    sel = os.environ.get("DIN_INV3_SPLIT", "none")
    if sel == "all":
        return 2
    if sel == "none":
        return 1
    return 2 if any(tok and tok in name for tok in sel.split(",")) else 1


class _BasicConv:
    def __init__(self, sd, name, stride=1, pad=(0, 0)):
        bn = {k: sd[f"{name}.bn.{k}"] for k in ("weight", "bias", "running_mean", "running_var")}
        bn["eps"] = 1e-3
        w, s, b = _fold_bn(sd[f"{name}.conv.weight"], bn)
        # split-weight mode is available (w_split = 2) but off by default: see _split_for
        self.conv = _Conv(w, b, s, stride=stride, pad=pad, relu=True, split=_split_for(name))
        self.c_in, self.c_out = w.shape[1], w.shape[0]

    def __call__(self, x, out=None, x_c_offset=0, y_c_offset=0):
        return self.conv(x, out=out, c_in=self.c_in, x_c_offset=x_c_offset, y_c_offset=y_c_offset)


MERGE_1X1 = os.environ.get("DIN_INV3_MERGE", "1") != "0"      # A/B knob: 0 = one launch per branch convolution


class _BranchHeads:
    """The 1x1 convolutions that open the branches of one Inception block, as ONE GEMM over the block input (weight rows
    stacked; ops.conv2d_branches_nhwc).  `first` (branch1x1) lands in the block's concat buffer, the others in a scratch
    slab the branches' next convolutions read as channel slices.  The pool branch's 1x1 runs BEFORE its average pool: both
    are linear and the pool's zero padding (count_include_pad) commutes with a bias-free 1x1, so the pool streams c_out
    instead of c_in channels (768 -> 192, 288 -> 64); its BatchNorm shift + ReLU follow the pool.  Column order puts every
    boundary the kernel needs (the y / y2 split, the no-ReLU range) on a multiple of 32."""

    def __init__(self, sd, first, others, pool, order):
        names = {"first": first, "pool": pool, **others}
        parts, self.off, col = [], {}, 0
        for key in order:
            bn = {k: sd[f"{names[key]}.bn.{k}"] for k in ("weight", "bias", "running_mean", "running_var")}
            bn["eps"] = 1e-3
            w, s, b = _fold_bn(sd[f"{names[key]}.conv.weight"], bn)
            if key == "pool":
                self.pool_bias, b = b.contiguous().float(), torch.zeros_like(b)
            parts.append((w, s, b))
            self.off[key] = (col, w.shape[0])
            col += w.shape[0]
        assert order[0] == "first"
        self.split_col = self.off["first"][1]
        self.norelu = (self.off["pool"][0], self.off["pool"][0] + self.off["pool"][1])
        assert self.split_col % 32 == 0 and self.norelu[0] % 32 == 0 and self.norelu[1] % 32 == 0
        self.c_in, self.c_total = parts[0][0].shape[1], col
        self.conv = _Conv(torch.cat([w for w, _, _ in parts]), torch.cat([b for _, _, b in parts]),
                          torch.cat([s for _, s, _ in parts]), relu=True, split=_split_for(first))

    def slice(self, key):
        """(channel offset, channels) of a branch inside the scratch slab."""
        o, c = self.off[key]
        return o - self.split_col, c

    def __call__(self, x, out):
        scratch = torch.empty(x.shape[:3] + (self.c_total - self.split_col,), dtype=torch.float16, device=x.device)
        ops.conv2d_branches_nhwc(x, self.conv._weight_for(x), self.conv.bias, out, scratch, split_col=self.split_col,
                                 norelu=self.norelu, c_in=self.c_in)
        return scratch

    def pool_tail(self, scratch, out, y_c_offset):
        o, c = self.slice("pool")
        ops.avgpool3_bias_relu_nhwc(scratch, self.pool_bias, out, c=c, x_c_offset=o, y_c_offset=y_c_offset)


class _InceptionA:   # Mixed_5b/5c/5d
    def __init__(self, sd, p):
        self.b5_2 = _BasicConv(sd, p + "branch5x5_2", pad=(2, 2))
        self.d2 = _BasicConv(sd, p + "branch3x3dbl_2", pad=(1, 1))
        self.d3 = _BasicConv(sd, p + "branch3x3dbl_3", pad=(1, 1))
        if MERGE_1X1:
            # [branch1x1 64 | 3x3dbl_1 64 | pool pf | 5x5_1 48]: split at 64, pool columns start at 128
            self.heads = _BranchHeads(sd, p + "branch1x1", {"d1": p + "branch3x3dbl_1", "s1": p + "branch5x5_1"},
                                      p + "branch_pool", ("first", "d1", "pool", "s1"))
            self.c_in, pf = self.heads.c_in, self.heads.off["pool"][1]
        else:
            self.b1 = _BasicConv(sd, p + "branch1x1")
            self.b5_1 = _BasicConv(sd, p + "branch5x5_1")
            self.d1 = _BasicConv(sd, p + "branch3x3dbl_1")
            self.bp = _BasicConv(sd, p + "branch_pool")
            self.c_in, pf = self.b1.c_in, self.bp.c_out
        self.c_out = 64 + 64 + 96 + pf

    def __call__(self, x, out):
        if MERGE_1X1:
            h = self.heads
            scratch = h(x, out)                                   # branch1x1 -> out[..., 0:64]
            o, _ = h.slice("s1")
            self.b5_2(scratch, out=out, x_c_offset=o, y_c_offset=64)
            o, _ = h.slice("d1")
            self.d3(self.d2(scratch, x_c_offset=o), out=out, y_c_offset=128)
            h.pool_tail(scratch, out, 224)
            return out
        self.b1(x, out=out, y_c_offset=0)
        self.b5_2(self.b5_1(x), out=out, y_c_offset=64)
        self.d3(self.d2(self.d1(x)), out=out, y_c_offset=128)
        self.bp(ops.avgpool2d_nhwc(x, 3, 1, 1, c=self.c_in), out=out, y_c_offset=224)
        return out


class _InceptionB:   # Mixed_6a
    def __init__(self, sd, p):
        self.b3 = _BasicConv(sd, p + "branch3x3", stride=2)
        self.d1 = _BasicConv(sd, p + "branch3x3dbl_1")
        self.d2 = _BasicConv(sd, p + "branch3x3dbl_2", pad=(1, 1))
        self.d3 = _BasicConv(sd, p + "branch3x3dbl_3", stride=2)
        self.c_in = self.b3.c_in

    def __call__(self, x, out):
        self.b3(x, out=out, y_c_offset=0)
        self.d3(self.d2(self.d1(x)), out=out, y_c_offset=384)
        ops.maxpool2d_nhwc(x, 3, 2, 0, out=out, c=self.c_in, y_c_offset=480)
        return out


class _InceptionC:   # Mixed_6b..6e
    def __init__(self, sd, p):
        if MERGE_1X1:
            # [branch1x1 192 | 7x7_1 c7 | 7x7dbl_1 c7 | pool 192]: split at 192, pool columns start at 192 + 2 c7
            self.heads = _BranchHeads(sd, p + "branch1x1", {"s1": p + "branch7x7_1", "d1": p + "branch7x7dbl_1"},
                                      p + "branch_pool", ("first", "s1", "d1", "pool"))
        else:
            self.b1 = _BasicConv(sd, p + "branch1x1")
            self.s1 = _BasicConv(sd, p + "branch7x7_1")
            self.d1 = _BasicConv(sd, p + "branch7x7dbl_1")
            self.bp = _BasicConv(sd, p + "branch_pool")
        self.s2 = _BasicConv(sd, p + "branch7x7_2", pad=(0, 3))
        self.s3 = _BasicConv(sd, p + "branch7x7_3", pad=(3, 0))
        self.d2 = _BasicConv(sd, p + "branch7x7dbl_2", pad=(3, 0))
        self.d3 = _BasicConv(sd, p + "branch7x7dbl_3", pad=(0, 3))
        self.d4 = _BasicConv(sd, p + "branch7x7dbl_4", pad=(3, 0))
        self.d5 = _BasicConv(sd, p + "branch7x7dbl_5", pad=(0, 3))

    def __call__(self, x, out):
        if MERGE_1X1:
            h = self.heads
            scratch = h(x, out)                                   # branch1x1 -> out[..., 0:192]
            o, _ = h.slice("s1")
            self.s3(self.s2(scratch, x_c_offset=o), out=out, y_c_offset=192)
            o, _ = h.slice("d1")
            self.d5(self.d4(self.d3(self.d2(scratch, x_c_offset=o))), out=out, y_c_offset=384)
            h.pool_tail(scratch, out, 576)
            return out
        self.b1(x, out=out, y_c_offset=0)
        self.s3(self.s2(self.s1(x)), out=out, y_c_offset=192)
        self.d5(self.d4(self.d3(self.d2(self.d1(x)))), out=out, y_c_offset=384)
        self.bp(ops.avgpool2d_nhwc(x, 3, 1, 1), out=out, y_c_offset=576)
        return out


class Inv3Plan:
    def __init__(self, sd, prefix="backbone."):
        bn = {k: sd[f"{prefix}Conv2d_1a_3x3.bn.{k}"] for k in ("weight", "bias", "running_mean", "running_var")}
        bn["eps"] = 1e-3
        w, s, b = _fold_bn(sd[prefix + "Conv2d_1a_3x3.conv.weight"], bn)
        self.stem = _Stem(w, b, scale=s, stride=2, pad=0)
        self.c2a = _BasicConv(sd, prefix + "Conv2d_2a_3x3")
        self.c2b = _BasicConv(sd, prefix + "Conv2d_2b_3x3", pad=(1, 1))
        self.c3b = _BasicConv(sd, prefix + "Conv2d_3b_1x1")
        self.c4a = _BasicConv(sd, prefix + "Conv2d_4a_3x3")
        self.m5b = _InceptionA(sd, prefix + "Mixed_5b.")
        self.m5c = _InceptionA(sd, prefix + "Mixed_5c.")
        self.m5d = _InceptionA(sd, prefix + "Mixed_5d.")
        self.m6a = _InceptionB(sd, prefix + "Mixed_6a.")
        self.m6 = [_InceptionC(sd, prefix + f"Mixed_6{c}.") for c in "bcde"]
        self.last_out1 = None

    def out_shape(self, h, w):
        h, w = _c(h, 3, 2), _c(w, 3, 2)      # 1a
        h, w = _c(h, 3, 1), _c(w, 3, 1)      # 2a   (2b keeps the size)
        h, w = _c(h, 3, 2), _c(w, 3, 2)      # max-pool
        h, w = _c(h, 3, 1), _c(w, 3, 1)      # 4a   (3b is 1x1)
        h, w = _c(h, 3, 2), _c(w, 3, 2)      # max-pool
        return h, w, D_STRIDE

    def __call__(self, images, out=None):
        """raw frames (fp32 NCHW or uint8 NHWC) -> multiscale map [F, OH, OW, 1088] (channels [0,1056) valid,
        rest untouched)."""
        F_ = images.shape[0]
        H, W = images.shape[1:3] if images.dtype == torch.uint8 else images.shape[2:4]
        oh, ow, _ = self.out_shape(H, W)
        dev = images.device
        if out is None:
            out = torch.zeros((F_, oh, ow, D_STRIDE), dtype=torch.float16, device=dev)
        x = self.stem(images)
        x = self.c2b(self.c2a(x))
        x = ops.maxpool2d_nhwc(x, 3, 2, 0)
        x = self.c4a(self.c3b(x))
        x = ops.maxpool2d_nhwc(x, 3, 2, 0)                                        # [F, oh, ow, 192]
        e = lambda c, hh=oh, ww=ow: torch.empty((F_, hh, ww, c), dtype=torch.float16, device=dev)  # noqa: E731
        x = self.m5b(x, e(256))
        x = self.m5c(x, e(288))
        self.m5d(x, out)                                                          # channels [0, 288) of the map
        h2, w2 = _c(oh, 3, 2), _c(ow, 3, 2)
        # Mixed_6a reads channels [0, 288) of the map in place (c_in = 288 < its channel stride)
        y = self.m6a(out, e(768, h2, w2))
        for blk in self.m6:
            y = blk(y, e(768, h2, w2))
        self.last_out1 = y
        ops.upsample_bilinear_nhwc(y, oh, ow, out=out, c=768, y_c_offset=288)    # F.interpolate + cat
        return out
