"""inception.py — the truncated Inception-v3 backbone plan (reference backbone/backbone.py:10-85) plus the
multiscale build of infer_model.py:165-172, on NHWC fp16.

torchvision's BasicConv2d = Conv2d(bias=False) + BatchNorm2d(eps=1e-3) + ReLU: BN (eval statistics) is
folded into the packed weights' scale and the epilogue bias.  Every branch convolution writes its output
channels directly at its offset of the block's concat buffer (`y_c_offset` / `y_c_stride` of the conv ABI),
so `torch.cat` never runs; Mixed_5d writes into channels [0, 288) of the multiscale map and Mixed_6e's
768 channels are bilinearly upsampled (align_corners=True) into channels [288, 1056) of the same map.
Channel counts that are not multiples of 64 (32, 48, 80, 96, 160, 288) need no padded buffers: the conv's
TMA tensor map is given the real channel extent and zero-fills the rest of the last 64-channel K block.

Training (BatchNorm in eval mode = folded, gamma / beta still train, as cfg.set_bn_eval leaves it):
`forward_train` keeps every activation, `backward` walks the graph in reverse with the same kernels the VGG-16 / ResNet-18
backward uses -- data gradients on the forward tcgen05 kernel (ReLU backward of the layer below fused into its epilogue),
weight gradients on conv_wgrad_tcgen05.cu (1x1 / 3x3 / 5x5 / 1x7 / 7x1), stride-2 convolutions through zero insertion --
plus the concat-slice ReLU mask, the pad-0 3x3/2 max-pool routing and the bilinear-resize adjoint (csrc/backbone_bwd.cu).
A block's input gradient is ONE data-gradient GEMM over the merged branch-head slab, not a sum of four.
"""
from __future__ import annotations

import os

import torch

from . import ops
from .engine import _Conv, _Stem, _fold_bn, _pack_dgrad_filters

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
        self.conv = _Conv(w, b, s, stride=stride, pad=pad, relu=True, split=_split_for(name), bn=bn)
        self.name, self.stride, self.pad = name, stride, pad
        self.c_in, self.c_out = w.shape[1], w.shape[0]
        self.kh, self.kw = w.shape[2], w.shape[3]

    def __call__(self, x, out=None, x_c_offset=0, y_c_offset=0):
        return self.conv(x, out=out, c_in=self.c_in, x_c_offset=x_c_offset, y_c_offset=y_c_offset)

    # ---- training
    def new_acc(self, device):
        z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=device)  # noqa: E731
        return {"dw": z(self.c_out, self.kh, self.kw, self.c_in), "dbeta": z(self.c_out), "dgamma": z(self.c_out)}

    def wgrad(self, acc, x, dz, inv_scale, x_c_offset=0):
        """dz: gradient w.r.t. this conv's pre-ReLU output (dense [n, oh, ow, c_out]); x: its saved input."""
        if self.stride == 2:                                   # zero insertion: the stride-1 kernel on the input grid
            dz = ops.scatter2_nhwc(dz, x.shape[1] - self.kh + 1 + 2 * self.pad[0], x.shape[2] - self.kw + 1 + 2 * self.pad[1])
        ops.conv2d_wgrad_nhwc(x, dz, acc["dw"], acc["dbeta"], pad=self.pad, inv_scale=inv_scale, x_c_offset=x_c_offset)
        return dz                                              # (zero-inserted for stride 2: dgrad takes the same tensor)

    def dgrad(self, dz, out=None, y_c_offset=0, relu_mask=None, residual=None):
        """d(input) = conv(dz, rot180(W)^T) with padding k - 1 - pad (dz zero-inserted for stride 2); optional fused ReLU
        mask of the layer below (relu_mask: shaped like `out`) or accumulation (residual = out: in place)."""
        pad = (self.kh - 1 - self.pad[0], self.kw - 1 - self.pad[1])
        return ops.conv2d_nhwc(dz, self.conv._w_dgrad, None, stride=1, pad=pad, relu=False, c_in=self.c_out, out=out,
                               y_c_offset=y_c_offset, relu_mask=relu_mask, residual=residual)

    def export(self, acc, grads):
        _export_folded(self.name, self.conv.w_src, self.conv.bn, self.conv.bn_scale, acc["dw"], acc["dbeta"], acc["dgamma"],
                       grads)


def _export_folded(name, w_src, bn, scale, dw_khwc, dbeta, dgamma, grads):
    """Accumulators of a convolution with its eval-mode BatchNorm folded in -> the reference's parameter names:
    d(gamma) = invstd * (<W, dW_folded> - mean * d(beta)) (din_bn_fold_grads_f32), dW = scale * dW_folded, OIHW."""
    dw = dw_khwc.permute(0, 3, 1, 2).contiguous()
    ops.bn_fold_grads(w_src, dw, dbeta, bn["running_mean"], bn["running_var"], dgamma, eps=bn["eps"])
    ops.scale_rows(dw, scale)
    grads[name + ".conv.weight"] = dw
    grads[name + ".bn.weight"], grads[name + ".bn.bias"] = dgamma, dbeta


MERGE_1X1 = os.environ.get("DIN_INV3_MERGE", "1") != "0"      # A/B knob: 0 = one launch per branch convolution (inference)


class _BranchHeads:
    """The 1x1 convolutions that open the branches of one Inception block, as ONE GEMM over the block input (weight rows
    stacked; ops.conv2d_branches_nhwc).  `first` (branch1x1) lands in the block's concat buffer, the others in the slab
    `zm` [n, h, w, c_total] at their merged column offsets (the first split_col columns of zm stay unused in the forward;
    the backward's gradient slab has the same layout and uses them), where the branches' next convolutions read them as
    channel slices.  The pool branch's 1x1 runs BEFORE its average pool: both are linear and the pool's zero padding
    (count_include_pad) commutes with a bias-free 1x1, so the pool streams c_out instead of c_in channels (768 -> 192,
    288 -> 64); its BatchNorm shift + ReLU follow the pool.  Column order puts every boundary the kernel needs (the
    y / y2 split, the no-ReLU range) on a multiple of 32."""

    def __init__(self, sd, first, others, pool, order):
        self.names = {"first": first, "pool": pool, **others}
        self.order = order
        parts, self.off, self.bn, col = [], {}, {}, 0
        for key in order:
            name = self.names[key]
            bn = {k: sd[f"{name}.bn.{k}"] for k in ("weight", "bias", "running_mean", "running_var")}
            bn["eps"] = 1e-3
            self.bn[key] = bn
            w, s, b = _fold_bn(sd[f"{name}.conv.weight"], bn)
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
        self.parts = parts
        self.conv = _Conv(torch.cat([w for w, _, _ in parts]), torch.cat([b for _, _, b in parts]),
                          torch.cat([s for _, s, _ in parts]), relu=True, split=_split_for(first))

    def slice(self, key):
        """(channel offset, channels) of a branch inside the slab."""
        return self.off[key]

    def __call__(self, x, out, x_c_offset=0):
        zm = torch.empty(x.shape[:3] + (self.c_total,), dtype=torch.float16, device=x.device)
        ops.conv2d_branches_nhwc(x, self.conv._weight_for(x), self.conv.bias, out, zm, split_col=self.split_col,
                                 norelu=self.norelu, c_in=self.c_in, x_c_offset=x_c_offset, y2_c_offset=self.split_col)
        return zm

    def pool_tail(self, zm, out, y_c_offset):
        o, c = self.slice("pool")
        ops.avgpool3_bias_relu_nhwc(zm, self.pool_bias, out, c=c, x_c_offset=o, y_c_offset=y_c_offset)

    # ---- training
    def new_acc(self, device):
        z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=device)  # noqa: E731
        return {"dw": z(self.c_total, 1, 1, self.c_in), "dbeta": z(self.c_total), "dgamma": z(self.c_total),
                "pool_dbeta": z(self.off["pool"][1])}

    def backward(self, acc, x, x_c_offset, out, d_out, dzm, out_pool_offset, inv_scale, dx, dx_residual):
        """dzm: the gradient slab (zm's layout) with every branch's columns filled in EXCEPT `first` and `pool`, which are
        derived here from the block's concat gradient d_out; then the merged weight gradient and ONE data-gradient GEMM
        into dx[..., 0:c_in] (accumulating in place when dx_residual)."""
        c1 = self.split_col
        ops.relu_bwd_slice_nhwc(out, d_out, dzm, c=c1)                                  # branch1x1: columns [0, c1)
        o, c = self.off["pool"]
        gp = torch.empty(out.shape[:3] + (c,), dtype=torch.float16, device=out.device)
        ops.relu_bwd_slice_nhwc(out, d_out, gp, c=c, y_c_offset=out_pool_offset, dy_c_offset=out_pool_offset)
        ops.colsum_nhwc(gp, acc["pool_dbeta"], c=c, inv_scale=inv_scale)                # the shift is added after the pool
        ops.avgpool2d_nhwc(gp, 3, 1, 1, out=dzm, y_c_offset=o)                          # the average pool is self-adjoint
        ops.conv2d_wgrad_nhwc(x, dzm, acc["dw"], acc["dbeta"], pad=(0, 0), inv_scale=inv_scale, x_c_offset=x_c_offset)
        ops.conv2d_nhwc(dzm, self.conv._w_dgrad, None, stride=1, pad=(0, 0), relu=False, c_in=self.c_total, out=dx,
                        residual=dx if dx_residual else None)

    def export(self, acc, grads):
        dw = acc["dw"]
        for key, (w, s, _) in zip(self.order, self.parts):
            o, c = self.off[key]
            dbeta = acc["pool_dbeta"] if key == "pool" else acc["dbeta"][o:o + c].contiguous()
            _export_folded(self.names[key], w.contiguous(), self.bn[key], s, dw[o:o + c].contiguous(), dbeta,
                           acc["dgamma"][o:o + c].contiguous(), grads)


def _e(ref, c, h=None, w=None):
    return torch.empty((ref.shape[0], ref.shape[1] if h is None else h, ref.shape[2] if w is None else w, c),
                       dtype=torch.float16, device=ref.device)


def _slice_dz(out, d_out, off, c):
    """Pre-ReLU gradient of a branch whose ReLU output is channels [off, off + c) of the concat buffer -> dense."""
    dz = _e(out, c)
    ops.relu_bwd_slice_nhwc(out, d_out, dz, c=c, y_c_offset=off, dy_c_offset=off)
    return dz


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

    def __call__(self, x, out, saved=None):
        if MERGE_1X1:
            h = self.heads
            zm = h(x, out)                                        # branch1x1 -> out[..., 0:64]
            self.b5_2(zm, out=out, x_c_offset=h.slice("s1")[0], y_c_offset=64)
            a2 = self.d2(zm, x_c_offset=h.slice("d1")[0])
            self.d3(a2, out=out, y_c_offset=128)
            h.pool_tail(zm, out, 224)
            if saved is not None:
                saved.update(x=x, out=out, zm=zm, a2=a2)
            return out
        self.b1(x, out=out, y_c_offset=0)
        self.b5_2(self.b5_1(x), out=out, y_c_offset=64)
        self.d3(self.d2(self.d1(x)), out=out, y_c_offset=128)
        self.bp(ops.avgpool2d_nhwc(x, 3, 1, 1, c=self.c_in), out=out, y_c_offset=224)
        return out

    def convs(self):
        return [self.b5_2, self.d2, self.d3]

    def backward(self, acc, sv, d_out, inv_scale, dx, dx_residual=False):
        h, out, zm = self.heads, sv["out"], sv["zm"]
        dzm = torch.empty_like(zm)
        dz = _slice_dz(out, d_out, 64, 64)                                          # 5x5 branch
        o = h.slice("s1")[0]
        self.b5_2.wgrad(acc["b5_2"], zm, dz, inv_scale, x_c_offset=o)
        self.b5_2.dgrad(dz, out=dzm, y_c_offset=o, relu_mask=zm)
        dz = _slice_dz(out, d_out, 128, 96)                                         # double 3x3 branch
        self.d3.wgrad(acc["d3"], sv["a2"], dz, inv_scale)
        dz = self.d3.dgrad(dz, relu_mask=sv["a2"])
        o = h.slice("d1")[0]
        self.d2.wgrad(acc["d2"], zm, dz, inv_scale, x_c_offset=o)
        self.d2.dgrad(dz, out=dzm, y_c_offset=o, relu_mask=zm)
        h.backward(acc["heads"], sv["x"], 0, out, d_out, dzm, 224, inv_scale, dx, dx_residual)


class _InceptionB:   # Mixed_6a
    def __init__(self, sd, p):
        self.b3 = _BasicConv(sd, p + "branch3x3", stride=2)
        self.d1 = _BasicConv(sd, p + "branch3x3dbl_1")
        self.d2 = _BasicConv(sd, p + "branch3x3dbl_2", pad=(1, 1))
        self.d3 = _BasicConv(sd, p + "branch3x3dbl_3", stride=2)
        self.c_in = self.b3.c_in

    def __call__(self, x, out, saved=None):
        self.b3(x, out=out, y_c_offset=0)
        a1 = self.d1(x)
        a2 = self.d2(a1)
        self.d3(a2, out=out, y_c_offset=384)
        ops.maxpool2d_nhwc(x, 3, 2, 0, out=out, c=self.c_in, y_c_offset=480)
        if saved is not None:
            saved.update(x=x, out=out, a1=a1, a2=a2)
        return out

    def convs(self):
        return [self.b3, self.d1, self.d2, self.d3]

    def backward(self, acc, sv, d_out, inv_scale, dx):
        """dx: the gradient buffer of the block input (channels [0, 288) of the multiscale map's gradient, which already
        holds RoIAlign's share): all three branches ACCUMULATE into it."""
        x, out = sv["x"], sv["out"]
        ops.maxpool3s2_bwd_nhwc(x, d_out, dx, c=self.c_in, pad=0, dy_c_offset=480, accumulate=True)
        dz = _slice_dz(out, d_out, 0, 384)
        dzu = self.b3.wgrad(acc["b3"], x, dz, inv_scale)
        self.b3.dgrad(dzu, out=dx, residual=dx)
        dz = _slice_dz(out, d_out, 384, 96)
        dzu = self.d3.wgrad(acc["d3"], sv["a2"], dz, inv_scale)
        dz = self.d3.dgrad(dzu, relu_mask=sv["a2"])
        self.d2.wgrad(acc["d2"], sv["a1"], dz, inv_scale)
        dz = self.d2.dgrad(dz, relu_mask=sv["a1"])
        self.d1.wgrad(acc["d1"], x, dz, inv_scale)
        self.d1.dgrad(dz, out=dx, residual=dx)


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

    def __call__(self, x, out, saved=None):
        if MERGE_1X1:
            h = self.heads
            zm = h(x, out)                                        # branch1x1 -> out[..., 0:192]
            s2 = self.s2(zm, x_c_offset=h.slice("s1")[0])
            self.s3(s2, out=out, y_c_offset=192)
            d2 = self.d2(zm, x_c_offset=h.slice("d1")[0])
            d3 = self.d3(d2)
            d4 = self.d4(d3)
            self.d5(d4, out=out, y_c_offset=384)
            h.pool_tail(zm, out, 576)
            if saved is not None:
                saved.update(x=x, out=out, zm=zm, s2=s2, d2=d2, d3=d3, d4=d4)
            return out
        self.b1(x, out=out, y_c_offset=0)
        self.s3(self.s2(self.s1(x)), out=out, y_c_offset=192)
        self.d5(self.d4(self.d3(self.d2(self.d1(x)))), out=out, y_c_offset=384)
        self.bp(ops.avgpool2d_nhwc(x, 3, 1, 1), out=out, y_c_offset=576)
        return out

    def convs(self):
        return [self.s2, self.s3, self.d2, self.d3, self.d4, self.d5]

    def backward(self, acc, sv, d_out, inv_scale, dx, dx_residual=False):
        h, out, zm = self.heads, sv["out"], sv["zm"]
        dzm = torch.empty_like(zm)
        dz = _slice_dz(out, d_out, 192, 192)                                        # 7x7 branch
        self.s3.wgrad(acc["s3"], sv["s2"], dz, inv_scale)
        dz = self.s3.dgrad(dz, relu_mask=sv["s2"])
        o = h.slice("s1")[0]
        self.s2.wgrad(acc["s2"], zm, dz, inv_scale, x_c_offset=o)
        self.s2.dgrad(dz, out=dzm, y_c_offset=o, relu_mask=zm)
        dz = _slice_dz(out, d_out, 384, 192)                                        # double 7x7 branch
        for name, below in (("d5", "d4"), ("d4", "d3"), ("d3", "d2")):
            layer = getattr(self, name)
            layer.wgrad(acc[name], sv[below], dz, inv_scale)
            dz = layer.dgrad(dz, relu_mask=sv[below])
        o = h.slice("d1")[0]
        self.d2.wgrad(acc["d2"], zm, dz, inv_scale, x_c_offset=o)
        self.d2.dgrad(dz, out=dzm, y_c_offset=o, relu_mask=zm)
        h.backward(acc["heads"], sv["x"], 0, out, d_out, dzm, 576, inv_scale, dx, dx_residual)


def _block_acc(blk, device):
    acc = {}
    if hasattr(blk, "heads"):
        acc["heads"] = blk.heads.new_acc(device)
    for name, layer in vars(blk).items():
        if isinstance(layer, _BasicConv):
            acc[name] = layer.new_acc(device)
    return acc


def _block_export(blk, acc, grads):
    if hasattr(blk, "heads"):
        blk.heads.export(acc["heads"], grads)
    for name, layer in vars(blk).items():
        if isinstance(layer, _BasicConv):
            layer.export(acc[name], grads)


class Inv3Plan:
    def __init__(self, sd, prefix="backbone."):
        bn = {k: sd[f"{prefix}Conv2d_1a_3x3.bn.{k}"] for k in ("weight", "bias", "running_mean", "running_var")}
        bn["eps"] = 1e-3
        w, s, b = _fold_bn(sd[prefix + "Conv2d_1a_3x3.conv.weight"], bn)
        self.prefix = prefix
        self.stem = _Stem(w, b, scale=s, stride=2, pad=0, bn=bn)
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

    def __call__(self, images, out=None, saved=None):
        """raw frames (fp32 NCHW or uint8 NHWC) -> multiscale map [F, OH, OW, 1088] (channels [0,1056) valid,
        rest untouched).  saved: a dict that receives every activation the backward needs (forward_train)."""
        F_ = images.shape[0]
        H, W = images.shape[1:3] if images.dtype == torch.uint8 else images.shape[2:4]
        oh, ow, _ = self.out_shape(H, W)
        dev = images.device
        if out is None:
            out = torch.zeros((F_, oh, ow, D_STRIDE), dtype=torch.float16, device=dev)
        sv = (lambda: {}) if saved is not None else (lambda: None)
        y1a = self.stem(images)
        y2a = self.c2a(y1a)
        y2b = self.c2b(y2a)
        p1 = ops.maxpool2d_nhwc(y2b, 3, 2, 0)
        y3b = self.c3b(p1)
        y4a = self.c4a(y3b)
        x = ops.maxpool2d_nhwc(y4a, 3, 2, 0)                                      # [F, oh, ow, 192]
        e = lambda c, hh=oh, ww=ow: torch.empty((F_, hh, ww, c), dtype=torch.float16, device=dev)  # noqa: E731
        s5b, s5c, s5d, s6a = sv(), sv(), sv(), sv()
        x = self.m5b(x, e(256), s5b)
        x = self.m5c(x, e(288), s5c)
        self.m5d(x, out, s5d)                                                     # channels [0, 288) of the map
        h2, w2 = _c(oh, 3, 2), _c(ow, 3, 2)
        # Mixed_6a reads channels [0, 288) of the map in place (c_in = 288 < its channel stride)
        y = self.m6a(out, e(768, h2, w2), s6a)
        s6 = []
        for blk in self.m6:
            s6.append(sv())
            y = blk(y, e(768, h2, w2), s6[-1])
        self.last_out1 = y
        ops.upsample_bilinear_nhwc(y, oh, ow, out=out, c=768, y_c_offset=288)    # F.interpolate + cat
        if saved is not None:
            saved.update(images=images, y1a=y1a, y2a=y2a, y2b=y2b, p1=p1, y3b=y3b, y4a=y4a, m5b=s5b, m5c=s5c, m5d=s5d,
                         m6a=s6a, m6=s6, hw2=(h2, w2))
        return out

    # -- training (BatchNorm in eval mode, i.e. folded, as cfg.set_bn_eval leaves it; gamma / beta still train) ----------
    def _all_convs(self):
        convs = [self.c2a, self.c2b, self.c3b, self.c4a]
        for blk in [self.m5b, self.m5c, self.m5d, self.m6a] + self.m6:
            convs += blk.convs()
        return [c.conv for c in convs] + [b.heads.conv for b in [self.m5b, self.m5c, self.m5d] + self.m6]

    def forward_train(self, images, out=None):
        assert MERGE_1X1, "the Inception-v3 backward is written for the merged branch heads (unset DIN_INV3_MERGE=0)"
        _pack_dgrad_filters(self._all_convs())
        saved = {}
        return self(images, out=out, saved=saved), saved

    def new_grads(self, device):
        z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=device)  # noqa: E731
        acc = {"stem": {"dw": z(32, 3, 3, 3), "dbeta": z(32), "dgamma": z(32)}}
        for name in ("c2a", "c2b", "c3b", "c4a"):
            acc[name] = getattr(self, name).new_acc(device)
        for name in ("m5b", "m5c", "m5d", "m6a"):
            acc[name] = _block_acc(getattr(self, name), device)
        acc["m6"] = [_block_acc(b, device) for b in self.m6]
        return acc

    def backward(self, sv, d_map, inv_scale, acc):
        """d_map: fp16 gradient (times the loss scale) w.r.t. this chunk's multiscale map [F, OH, OW, 1088]; modified in
        place (Mixed_6a's input gradient joins RoIAlign's in channels [0, 288)); accumulates into `acc`."""
        h2, w2 = sv["hw2"]
        d = ops.upsample_bilinear_bwd_nhwc(d_map, h2, w2, c=768, dy_c_offset=288)   # gradient of Mixed_6e's output
        for blk, a, s in reversed(list(zip(self.m6, acc["m6"], sv["m6"]))):
            dx = _e(d, 768)
            blk.backward(a, s, d, inv_scale, dx)
            d = dx
        self.m6a.backward(acc["m6a"], sv["m6a"], d, inv_scale, d_map)
        d = d_map                                                                # channels [0, 288): Mixed_5d's output
        for name, c_in in (("m5d", 288), ("m5c", 256), ("m5b", 192)):
            dx = _e(d, c_in)
            getattr(self, name).backward(acc[name], sv[name], d, inv_scale, dx)
            d = dx
        y4a, y3b, p1, y2b, y2a, y1a = (sv[k] for k in ("y4a", "y3b", "p1", "y2b", "y2a", "y1a"))
        dz = ops.maxpool3s2_bwd_nhwc(y4a, d, torch.empty_like(y4a), c=192, pad=0)   # + Conv2d_4a's ReLU backward
        self.c4a.wgrad(acc["c4a"], y3b, dz, inv_scale)
        dz = self.c4a.dgrad(dz, relu_mask=y3b)
        self.c3b.wgrad(acc["c3b"], p1, dz, inv_scale)
        d = self.c3b.dgrad(dz)
        dz = ops.maxpool3s2_bwd_nhwc(y2b, d, torch.empty_like(y2b), c=64, pad=0)
        self.c2b.wgrad(acc["c2b"], y2a, dz, inv_scale)
        dz = self.c2b.dgrad(dz, relu_mask=y2a)
        self.c2a.wgrad(acc["c2a"], y1a, dz, inv_scale)
        dz = self.c2a.dgrad(dz, relu_mask=y1a)
        a0 = acc["stem"]
        ops.stem_wgrad(sv["images"], dz, a0["dw"], a0["dbeta"], stride=2, pad=0, inv_scale=inv_scale, prep=True)

    def export_grads(self, acc, grads):
        a0 = acc["stem"]
        name = self.prefix + "Conv2d_1a_3x3"
        dw = a0["dw"].clone()                                                     # OIHW already
        ops.bn_fold_grads(self.stem.w_src, dw, a0["dbeta"], self.stem.bn["running_mean"], self.stem.bn["running_var"],
                          a0["dgamma"], eps=self.stem.bn["eps"])
        ops.scale_rows(dw, self.stem.bn_scale)
        grads[name + ".conv.weight"] = dw
        grads[name + ".bn.weight"], grads[name + ".bn.bias"] = a0["dgamma"], a0["dbeta"]
        for n in ("c2a", "c2b", "c3b", "c4a"):
            getattr(self, n).export(acc[n], grads)
        for n in ("m5b", "m5c", "m5d", "m6a"):
            _block_export(getattr(self, n), acc[n], grads)
        for blk, a in zip(self.m6, acc["m6"]):
            _block_export(blk, a, grads)
