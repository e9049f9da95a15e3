"""engine.py — the DIN stage-2 forward plan on one B200.

Host-side orchestration only: it packs the model's weights into the layouts the kernels want (once per
weight version), owns the activation workspaces, and issues the kernel sequence through the C ABI
(`ops.py`).  No arithmetic happens in Python/torch here.

Data layout in HBM
  images        fp32 NCHW [B*T, 3, H, W]        raw 0..255, as the reference's loader produces them, or
                uint8 NHWC [B*T, H, W, 3]       the decoded frame before the loader's transpose + float()
  activations   fp16 NHWC, two ping-pong slabs sized for the largest layer of one frame chunk
  feature map   fp16 NHWC [B*T, OH, OW, D]      all frames (0.9 MB / frame for VGG-16)
  crops         fp16 [B*T*N, 25, D]             RoIAlign output == A operand of the embedding GEMM
  person feats  fp32 [B, T, N, C]               everything after the embedding GEMM

Reference path restated (infer_model.py:141-234 / 1226-1319): prep -> backbone -> RoIAlign -> fc_emb_1 ->
nl_emb_1 -> ReLU -> (point_conv -> point_ln -> ReLU) -> DPI -> fuse + dpi_nl -> max_N -> fc_activities
-> mean_T.
"""
from __future__ import annotations

import os
import threading

import torch

from . import ops


def _fold_bn(conv_w, bn):
    """conv weight + BatchNorm (eval: running statistics) -> (weight, per-channel scale, bias)."""
    scale = bn["weight"] / torch.sqrt(bn["running_var"] + bn["eps"])
    bias = bn["bias"] - bn["running_mean"] * scale
    return conv_w, scale.contiguous(), bias.contiguous()


class _PackBatch:
    """`with _PackBatch():` -- every _Conv built inside registers its weight instead of packing it, and ONE
    din_pack_conv_weights_f16 launch packs them all on exit (a training step rebuilds the plan after every optimizer
    step: ~25 latency-bound packs back to back were 2.8 ms of a 43 ms VGG-16 step)."""
    active = None

    def __enter__(self):
        self.outer, self.convs = _PackBatch.active, []
        _PackBatch.active = self
        return self

    def __exit__(self, *exc):
        _PackBatch.active = self.outer
        if exc[0] is None:
            _pack_convs(self.convs)
        return False


def _pack_convs(convs):
    todo = [c for c in convs if c._w is None]
    for c, t in zip(todo, ops.pack_conv_weights([(c.w_src, c.bn_scale, c.split, False) for c in todo])):
        c._w = t


def _pack_dgrad_filters(convs):
    """The data-gradient filters of `convs` (packed straight from the forward weights, one launch)."""
    todo = [c for c in convs if c._w_dgrad is None]
    for c, t in zip(todo, ops.pack_conv_weights([(c.w_src, c.bn_scale, 1, True) for c in todo])):
        c._w_dgrad = t


# A launch with fewer output pixels than one wave of 128-pixel tiles (148 SMs) is latency-bound: the tensor pipe idles
# most of its ~10 us either way.  Such launches run with the weight split into hi + lo fp16 parts (w_split = 2: both
# multiplied with the same A tile into the same accumulator, i.e. weights exact to ~22 bits) -- precision that costs
# nothing exactly where the path needs it most: tiny inputs (one frame, one actor) have no averaging over pixels or
# actors to hide operand rounding behind (tests/test_edge_cases_gpu.py holds T = N = 1 to the same 1e-3 as every
# BASELINE shape).  Full-size launches are untouched.
# Frames per backbone launch.  A chunk bounds the activation workspace; beyond that, bigger is better (every launch is a
# persistent kernel that pays its prologue, pipeline fill and last-wave tail once): the inference forward sizes chunks
# so that the widest activation of a chunk stays under CHUNK_BYTES.  ResNet-18's late layers (23 x 40 maps) went from
# 470 to > 700 TFLOP/s when their launches grew from 16 to 64+ frames.  DIN_FRAMES_PER_CHUNK pins a fixed count.
CHUNK_BYTES = int(float(os.environ.get("DIN_CHUNK_GB", "4")) * (1 << 30))
# widest fp16 NHWC activation per frame, in bytes per input pixel: VGG-16 conv1_x (64 ch at full size), ResNet-18 stem
# (64 ch at 1/4 of the pixels), Inception-v3 Conv2d_2b (64 ch at ~1/4)
_ACT_BYTES_PER_PIXEL = {"vgg16": 128.0, "res18": 32.0, "inv3": 32.0}


def frames_per_chunk(backbone_name, n_frames, h, w, fixed=None):
    if fixed:
        return min(n_frames, fixed)
    per_frame = _ACT_BYTES_PER_PIXEL.get(backbone_name, 128.0) * h * w
    cap = max(1, int(CHUNK_BYTES // per_frame))
    n_chunks = -(-n_frames // cap)
    return -(-n_frames // n_chunks)                 # even split: 80 frames with cap 36 -> 27 + 27 + 26


FUSE_CONV1 = os.environ.get("DIN_FUSE_CONV1", "1") != "0"      # A/B knob: 0 = stand-alone stem + conv1_2
FUSE_STEM_POOL = os.environ.get("DIN_FUSE_STEM_POOL", "1") != "0"   # A/B knob: 0 = ResNet-18 stem and max-pool separately
SMALL_LAUNCH_PIXELS = 148 * 128 if os.environ.get("DIN_SMALL_EXACT", "1") != "0" else 0
# ... except the dense GEMM (n = h = 1: fc_emb_1) above 1 GFLOP: a full batch has only 960 'pixels' (actor rows) but
# K = 12 800 .. 27 200 -- 25-53 GFLOP, bound by weight traffic, where the second weight part doubled its time
# (0.27 -> 0.12 ms per step for VGG-16)
SMALL_LAUNCH_FLOPS = 1e9
# fewer actor rows than half an MMA tile: fp32 crops + fp32 fc_emb_1 (din_linear_f32)
SMALL_EMBED_ROWS = 64 if os.environ.get("DIN_SMALL_EMBED_F32", "1") != "0" else 0


def _plan_tensor(v, device):
    """A model tensor as the kernels want it: on the plan's device, fp32, and 16-byte aligned (the parameters of an
    nn.DataParallel replica are slices of one coalesced broadcast buffer and start at arbitrary 4-byte offsets)."""
    t = v.detach().to(device, torch.float32) if v.is_floating_point() else v.detach().to(device)
    if t.data_ptr() % 16 != 0:
        t = t.clone(memory_format=torch.contiguous_format)
    return t


class _InferenceFlag(threading.local):
    """True while DinEngine.features() (the no-grad inference forward) runs a backbone plan -- per THREAD: nn.DataParallel
    runs one replica per device in concurrent threads, and with a process-wide flag the replica that finished first switched
    the other one's remaining small launches from the exact (hi + lo) weights to the single-part ones (logits 5e-4 apart from
    the single-device result: tests/test_dropin_gpu.py::test_data_parallel_two_devices_one_plan_per_device)."""
    on = False


_INFERENCE = _InferenceFlag()


class _Conv:
    """One tcgen05 convolution of the plan (weights packed fp16 [co][kh][kw][ci], fp32 bias)."""

    def __init__(self, w, bias, scale=None, stride=1, pad=(0, 0), relu=True, pool2=False, split=1, bn=None):
        self.w_src = w                     # fp32 OIHW, kept by reference for the (lazily packed) dgrad filter
        self.bn_scale = scale              # folded eval-mode BatchNorm: gamma / sqrt(var + eps) (None: no BN)
        self.bn = bn                       # its {'weight', 'bias', ...} (training: d(gamma) needs gamma and beta)
        self.w_src = w = w.contiguous()
        self._w_dgrad = None
        self._w_exact = None               # hi + lo parts, packed on first use by a latency-bound launch
        self._w, self.split = None, split
        if _PackBatch.active is not None:
            _PackBatch.active.convs.append(self)
        else:
            _pack_convs([self])
        self.bias = None if bias is None else bias.contiguous().float()
        self.stride, self.pad, self.relu, self.pool2 = stride, pad, relu, pool2
        self.c_out, self.c_in = w.shape[0], w.shape[1]
        self.c_in_padded = (self.c_in + 63) // 64 * 64

    @property
    def w(self):
        assert self._w is not None, "weight used inside the _PackBatch block that defers its packing"
        return self._w

    def _weight_for(self, x):
        # inference plans only: a training step re-packs every weight after the optimizer step, and packing a second
        # (hi + lo) copy of the late layers each step cost ResNet-18's step 47 ms of packing kernels (measured)
        px = x.shape[0] * x.shape[1] * x.shape[2]
        if (not _INFERENCE.on or self.split == 2 or px > SMALL_LAUNCH_PIXELS * self.stride * self.stride
                or (x.shape[0] == 1 and x.shape[1] == 1 and 2.0 * px * self.w_src.numel() > SMALL_LAUNCH_FLOPS)):
            return self.w
        if self._w_exact is None:
            self._w_exact = ops.pack_conv_weights([(self.w_src, self.bn_scale, 2, False)])[0]
        return self._w_exact

    def __call__(self, x, out=None, residual=None, **kw):
        kw.setdefault("c_in", self.c_in)
        return ops.conv2d_nhwc(x, self._weight_for(x), self.bias, stride=self.stride, pad=self.pad, relu=self.relu,
                               residual=residual, out=out, pool2=self.pool2, **kw)

    def dgrad(self, dz, relu_mask=None):
        """dX of a stride-1 convolution (3x3 pad 1, or 1x1) = the forward kernel on dZ with the filter rotated by 180
        degrees and its channel axes swapped (packed once per weight version; a folded BN scale is multiplied in: one-time
        weight prep).  Stride-2 layers pass the zero-inserted dZ (ops.scatter2_nhwc)."""
        k = self.w_src.shape[2]
        assert k == self.w_src.shape[3] and k in (1, 3)
        _pack_dgrad_filters([self])                  # normally done for the whole plan by forward_train
        return ops.conv2d_nhwc(dz, self._w_dgrad, None, stride=1, pad=(k // 2, k // 2), relu=False, relu_mask=relu_mask)


class _Stem:
    def __init__(self, w, bias, scale=None, stride=1, pad=0, bn=None, relu=True):
        self.bn_scale, self.bn, self.relu = scale, bn, relu
        self.w_src = w.contiguous().float()       # un-folded (d(gamma) of a folded BatchNorm needs it)
        if scale is not None:
            w = w * scale.view(-1, 1, 1, 1)       # one-time weight prep (BN folding), not on the hot path
        self.w, self.bias, self.stride, self.pad = w.contiguous().float(), bias.contiguous().float(), stride, pad

    def __call__(self, images):
        return ops.stem_conv(images, self.w, self.bias, stride=self.stride, pad=self.pad, relu=self.relu, prep=True)


# ------------------------------------------------------------------------------------------------
# backbones: each returns the NHWC fp16 feature map [frames, OH, OW, D] for a chunk of frames
# ------------------------------------------------------------------------------------------------
class VGG16Plan:
    """torchvision vgg16().features (backbone.py:88-99): 13 x (3x3 conv + ReLU), 5 x maxpool 2x2."""
    CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"]

    def __init__(self, sd, prefix="backbone.features."):
        self.layers, self.param_names = [], []
        idx = 0
        for i, v in enumerate(self.CFG):
            if v == "M":
                idx += 1        # the 2x2 max-pool is fused into the preceding conv's epilogue
                continue
            w, b = sd[f"{prefix}{idx}.weight"], sd[f"{prefix}{idx}.bias"]
            pool = i + 1 < len(self.CFG) and self.CFG[i + 1] == "M"
            if not self.layers:
                self.layers.append(_Stem(w, b, stride=1, pad=1))
            else:
                self.layers.append(_Conv(w, b, stride=1, pad=(1, 1), relu=True, pool2=pool))
            self.param_names.append(f"{prefix}{idx}")
            idx += 2
        self.out_channels = 512

    # -- training: forward that keeps what the backward needs, and the backward itself
    def forward_train(self, images, out=None):
        """As __call__, but every conv's ReLU output is kept BEFORE its max-pool (the pool runs as its own kernel;
        max over the same fp16 values as the fused epilogue, so the result is bit-identical).
        -> (feature map, saved = [(conv input, ReLU output, pooled?)] per conv)."""
        saved = []
        x = images
        last = len(self.layers) - 1
        _pack_dgrad_filters(self.layers[1:])
        for i, layer in enumerate(self.layers):
            if i == 0:
                y, pooled = layer(x), False
                a = y
            else:
                pooled = layer.pool2
                y = ops.conv2d_nhwc(x, layer.w, layer.bias, stride=1, pad=(1, 1), relu=True, pool2=False, c_in=layer.c_in)
                a = ops.maxpool2d_nhwc(y, 2, 2, 0, out=out if i == last else None) if pooled else y
            saved.append((x, y, pooled))
            x = a
        return x, saved

    def new_grads(self, device):
        """Zero-filled fp32 accumulators: the stem's in OIHW, the others in the kernels' [co][kh][kw][ci] order."""
        acc = []
        for i, layer in enumerate(self.layers):
            co = layer.w.shape[0]
            shape = (co, 3, 3, 3) if i == 0 else (co, 3, 3, layer.c_in)
            acc.append((torch.zeros(shape, dtype=torch.float32, device=device),
                        torch.zeros((co,), dtype=torch.float32, device=device)))
        return acc

    def backward(self, saved, d_out, inv_scale, acc):
        """d_out: fp16 gradient (times the loss scale) w.r.t. this chunk's feature map; accumulates into `acc`."""
        d, masked = d_out, False
        for i in reversed(range(len(self.layers))):
            x_in, y, pooled = saved[i]
            dz = d if masked else ops.relu_pool_bwd_nhwc(y, d, pooled)
            dw, db = acc[i]
            if i == 0:
                ops.stem_wgrad(x_in, dz, dw, db, inv_scale=inv_scale, prep=True)
            else:
                ops.conv2d_wgrad_nhwc(x_in, dz, dw, db, pad=(1, 1), inv_scale=inv_scale)
                # the layer below feeds this one without a pool: its ReLU backward rides in this dgrad's epilogue
                masked = not saved[i - 1][2]
                d = self.layers[i].dgrad(dz, relu_mask=saved[i - 1][1] if masked else None)

    def export_grads(self, acc, grads):
        """accumulators -> {reference parameter name: OIHW gradient} (layout only)."""
        for i, (dw, db) in enumerate(acc):
            name = self.param_names[i]
            grads[name + ".weight"] = dw if i == 0 else dw.permute(0, 3, 1, 2).contiguous()
            grads[name + ".bias"] = db

    def __call__(self, images, out=None):
        x = images
        last = len(self.layers) - 1
        first = 0
        stem, c12 = self.layers[0], self.layers[1]
        n_px = images.shape[0] * (images.shape[1] * images.shape[2] if images.dtype == torch.uint8
                                  else images.shape[2] * images.shape[3])
        if FUSE_CONV1 and ops.stem_pair_supported(images) and n_px > SMALL_LAUNCH_PIXELS:
            # conv1_1 + conv1_2 (+ pool) in one launch: conv1_1's output never reaches HBM (inference only; the
            # training forward keeps every activation).  Tiny launches stay on the two-kernel path (exact weights).
            x = ops.stem_conv_pair(images, stem.w, stem.bias, c12.w, c12.bias, relu=True, pool2=c12.pool2, prep=True)
            first = 2
        for i in range(first, len(self.layers)):
            layer = self.layers[i]
            x = layer(x, out=out) if i == last else layer(x)
        return x

    def out_shape(self, h, w):
        return h // 32, w // 32, 512


class Res18Plan:
    """torchvision resnet18 up to layer4 (backbone.py:115-132), eval-mode BN folded into the convs."""

    def __init__(self, sd, prefix="backbone.features.", bn_train=False):
        """bn_train: BatchNorm on batch statistics (module.train() without cfg.set_bn_eval): the convolutions stay
        un-folded and every BatchNorm runs as its own kernels (csrc/bn_train.cu); forward_train / backward only."""
        self.bn_train = bn_train
        self._nbt = []

        def bn(p):
            t = sd.get(f"{p}.num_batches_tracked")
            if t is not None and all(t is not u for u in self._nbt):   # bn() runs twice per layer; a duplicate in the
                self._nbt.append(t)                                    # _foreach_add_ list is a read-modify-write race
            return {k: sd[f"{p}.{k}"] for k in ("weight", "bias", "running_mean", "running_var")} | {"eps": 1e-5}

        def _fold(w, b):
            if bn_train:
                return w, None, None
            return _fold_bn(w, b)

        w, s, b = _fold(sd[prefix + "0.weight"], bn(prefix + "1"))
        if bn_train:
            b = torch.zeros(w.shape[0], dtype=torch.float32, device=w.device)
        self.stem = _Stem(w, b, scale=s, stride=2, pad=3, bn=bn(prefix + "1"), relu=not bn_train)
        self.prefix = prefix
        self.block_names = []
        self.blocks = []
        chans = [64, 128, 256, 512]
        for li, c in enumerate(chans):
            for bi in range(2):
                p = f"{prefix}{4 + li}.{bi}."
                stride = 2 if (li > 0 and bi == 0) else 1
                w1, s1, b1 = _fold(sd[p + "conv1.weight"], bn(p + "bn1"))
                w2, s2, b2 = _fold(sd[p + "conv2.weight"], bn(p + "bn2"))
                conv1 = _Conv(w1, b1, s1, stride=stride, pad=(1, 1), relu=not bn_train, bn=bn(p + "bn1"))
                conv2 = _Conv(w2, b2, s2, stride=1, pad=(1, 1), relu=not bn_train, bn=bn(p + "bn2"))   # ReLU after the residual add
                down = None
                if (p + "downsample.0.weight") in sd:
                    wd, sdn, bd = _fold(sd[p + "downsample.0.weight"], bn(p + "downsample.1"))
                    down = _Conv(wd, bd, sdn, stride=stride, pad=(0, 0), relu=False, bn=bn(p + "downsample.1"))
                self.blocks.append((conv1, conv2, down))
                self.block_names.append(p)
        self.out_channels = 512

    def __call__(self, images, out=None):
        assert not self.bn_train, "a batch-statistics plan only runs forward_train (DinEngine.features does that)"
        if FUSE_STEM_POOL and self.stem.relu and ops.stem_pool_supported(images):
            # conv1 + bn1 + relu + maxpool in one launch: the half-resolution 64-channel map never reaches HBM (inference;
            # the training forward keeps it for the backward)
            x = ops.stem_conv_pool(images, self.stem.w, self.stem.bias, prep=True)
        else:
            x = self.stem(images)
            x = ops.maxpool2d_nhwc(x, 3, 2, 1)
        last = len(self.blocks) - 1
        for i, (conv1, conv2, down) in enumerate(self.blocks):
            identity = x if down is None else down(x)
            y = conv1(x)
            x = conv2(y, residual=identity, out=out if i == last else None)
        return x

    def out_shape(self, h, w):
        def c(v, k, s, p):
            return (v + 2 * p - k) // s + 1
        h, w = c(h, 7, 2, 3), c(w, 7, 2, 3)
        for _ in range(4):
            h, w = c(h, 3, 2, 1), c(w, 3, 2, 1)
        return h, w, 512

    # -- training (BatchNorm in eval mode, i.e. folded, as cfg.set_bn_eval leaves it) -----------
    def forward_train(self, images, out=None):
        """As __call__, keeping the stem output and every block's (input, conv1 output, identity, output)."""
        if self.bn_train:
            return self._forward_train_bn(images, out)
        _pack_dgrad_filters([c for convs in self.blocks for c in convs if c is not None])
        x0 = self.stem(images)
        x = ops.maxpool2d_nhwc(x0, 3, 2, 1)
        saved = {"images": images, "x0": x0, "blocks": []}
        last = len(self.blocks) - 1
        for i, (conv1, conv2, down) in enumerate(self.blocks):
            identity = x if down is None else down(x)
            a1 = conv1(x)
            y = conv2(a1, residual=identity, out=out if i == last else None)
            saved["blocks"].append((x, a1, identity, y))
            x = y
        return x, saved

    @staticmethod
    def _acc(co, ci, device, taps=3):
        z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=device)  # noqa: E731
        return {"dw": z(co, taps, taps, ci), "dbeta": z(co), "dgamma": z(co)}

    def new_grads(self, device):
        """Zero-filled fp32 accumulators of the FOLDED convolutions ([co][3][3][ci]; the stem's in OIHW; the 1x1
        shortcut's as the centre tap of a 3x3) plus d(beta) / d(gamma) of every BatchNorm."""
        z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=device)  # noqa: E731
        acc = {"stem": {"dw": z(64, 3, 7, 7), "dbeta": z(64), "dgamma": z(64)}, "blocks": []}
        for conv1, conv2, down in self.blocks:
            acc["blocks"].append((self._acc(conv1.c_out, conv1.c_in, device), self._acc(conv2.c_out, conv2.c_in, device),
                                  None if down is None else self._acc(down.c_out, down.c_in, device)))
        return acc

    def backward(self, saved, d_out, inv_scale, acc):
        """d_out: fp16 gradient (times the loss scale) w.r.t. this chunk's feature map; accumulates into `acc`.
        Stride-2 layers go through zero insertion (ops.scatter2_nhwc), so dgrad / wgrad stay the stride-1 kernels."""
        if self.bn_train:
            return self._backward_bn(saved, d_out, inv_scale, acc)
        d = d_out
        for bi in reversed(range(len(self.blocks))):
            x_in, a1, identity, y = saved["blocks"][bi]
            conv1, conv2, down = self.blocks[bi]
            a_c1, a_c2, a_dn = acc["blocks"][bi]
            h_in, w_in = x_in.shape[1:3]
            dz2 = ops.relu_pool_bwd_nhwc(y, d, False)                         # ReLU after the residual add
            ops.conv2d_wgrad_nhwc(a1, dz2, a_c2["dw"], a_c2["dbeta"], pad=(1, 1), inv_scale=inv_scale)
            dz1 = conv2.dgrad(dz2, relu_mask=a1)                              # conv1's ReLU backward in the epilogue
            dz1u = ops.scatter2_nhwc(dz1, h_in, w_in) if conv1.stride == 2 else dz1
            ops.conv2d_wgrad_nhwc(x_in, dz1u, a_c1["dw"], a_c1["dbeta"], pad=(1, 1), inv_scale=inv_scale)
            dx = conv1.dgrad(dz1u)
            if down is None:
                dx = ops.add_f16(dx, dz2)                                     # the identity shortcut
            else:
                # 1x1 stride-2 shortcut: its weight gradient is the centre tap of a 3x3 pad-1 wgrad on the zero-inserted
                # dZ (9x the useful work on three small layers, no extra tensor-core kernel)
                dz2u = ops.scatter2_nhwc(dz2, h_in, w_in)
                ops.conv2d_wgrad_nhwc(x_in, dz2u, a_dn["dw"], a_dn["dbeta"], pad=(1, 1), inv_scale=inv_scale)
                ops.scatter2_nhwc(down.dgrad(dz2), h_in, w_in, dst=dx)
            d = dx
        dz0 = ops.maxpool3s2_relu_bwd_nhwc(saved["x0"], d)
        a0 = acc["stem"]
        ops.stem_wgrad(saved["images"], dz0, a0["dw"], a0["dbeta"], stride=2, pad=3, inv_scale=inv_scale, prep=True)
        # d(gamma) of every folded BatchNorm comes from dW and d(beta) at export time (din_bn_fold_grads_f32)

    # -- BatchNorm on batch statistics (scripts/train_collective_stage2_dynamic.py: ResNet-18 trained with
    #    cfg.set_bn_eval = False): conv -> raw z -> stats -> normalise (+ residual) -> ReLU, all frames of the step at once
    def _bn(self, layer, z, residual=None, relu=True, out=None):
        b = layer.bn
        y, stats = ops.bn_train_forward(z, b["weight"], b["bias"], b["running_mean"], b["running_var"], eps=b["eps"],
                                        residual=residual, relu=relu, out=out)
        return y, stats

    def _forward_train_bn(self, images, out=None):
        convs = [c for cs in self.blocks for c in cs if c is not None]
        _pack_dgrad_filters(convs)
        z0 = self.stem(images)
        x0, st0 = self._bn(self.stem, z0)
        x = ops.maxpool2d_nhwc(x0, 3, 2, 1)
        saved = {"images": images, "z0": z0, "x0": x0, "st0": st0, "blocks": []}
        last = len(self.blocks) - 1
        for i, (conv1, conv2, down) in enumerate(self.blocks):
            zd = std = None
            identity = x
            # raw convolution outputs in fp32: the normalised activation is then rounded to fp16 once, as the folded
            # eval-mode epilogue does (fp16 z: twice the ReLU decision flips, gradients 15-20 % off instead of 2-7 %)
            if down is not None:
                zd = down(x, out_f32=True)
                identity, std = self._bn(down, zd, relu=False)
            z1 = conv1(x, out_f32=True)
            a1, st1 = self._bn(conv1, z1)
            z2 = conv2(a1, out_f32=True)
            y, st2 = self._bn(conv2, z2, residual=identity, out=out if i == last else None)
            saved["blocks"].append((x, z1, a1, st1, z2, y, st2, zd, std))
            x = y
        if self._nbt:
            torch._foreach_add_(self._nbt, 1)                # num_batches_tracked
        for c in [self.stem] + convs:                        # the kernels wrote the running statistics through raw
            for k in ("running_mean", "running_var"):        # pointers: tell torch (plan caches key on ._version)
                torch.autograd.graph.increment_version(c.bn[k])
        return x, saved

    def _backward_bn(self, saved, d_out, inv_scale, acc):
        def bn_bwd(layer, a, g, z, stats):
            return ops.bn_train_backward(g, z, stats[0], stats[1], layer.bn["weight"], a["dbeta"], a["dgamma"],
                                         inv_scale=inv_scale)

        d = d_out
        for bi in reversed(range(len(self.blocks))):
            x_in, z1, a1, st1, z2, y, st2, zd, std = saved["blocks"][bi]
            conv1, conv2, down = self.blocks[bi]
            a_c1, a_c2, a_dn = acc["blocks"][bi]
            h_in, w_in = x_in.shape[1:3]
            g2 = ops.relu_pool_bwd_nhwc(y, d, False)                          # also the shortcut branch's gradient
            dz2 = bn_bwd(conv2, a_c2, g2, z2, st2)
            ops.conv2d_wgrad_nhwc(a1, dz2, a_c2["dw"], None, pad=(1, 1), inv_scale=inv_scale)
            dz1 = bn_bwd(conv1, a_c1, conv2.dgrad(dz2, relu_mask=a1), z1, st1)
            dz1u = ops.scatter2_nhwc(dz1, h_in, w_in) if conv1.stride == 2 else dz1
            ops.conv2d_wgrad_nhwc(x_in, dz1u, a_c1["dw"], None, pad=(1, 1), inv_scale=inv_scale)
            dx = conv1.dgrad(dz1u)
            if down is None:
                dx = ops.add_f16(dx, g2)
            else:
                dzd = bn_bwd(down, a_dn, g2, zd, std)
                ops.conv2d_wgrad_nhwc(x_in, ops.scatter2_nhwc(dzd, h_in, w_in), a_dn["dw"], None, pad=(1, 1),
                                      inv_scale=inv_scale)
                ops.scatter2_nhwc(down.dgrad(dzd), h_in, w_in, dst=dx)
            d = dx
        a0 = acc["stem"]
        dz0 = bn_bwd(self.stem, a0, ops.maxpool3s2_relu_bwd_nhwc(saved["x0"], d), saved["z0"], saved["st0"])
        ops.stem_wgrad(saved["images"], dz0, a0["dw"], torch.zeros_like(a0["dbeta"]), stride=2, pad=3,
                       inv_scale=inv_scale, prep=True)

    def export_grads(self, acc, grads):
        """accumulators -> {reference parameter name: gradient}: un-fold the BN scale from the conv weight gradients
        (din_scale_rows_f32), OIHW layout (permutes only)."""
        def put(conv_name, bn_name, conv, a, w_oihw):
            w_oihw = w_oihw.contiguous()
            if w_oihw.data_ptr() == a["dw"].data_ptr():
                w_oihw = w_oihw.clone()         # the stem's accumulator is OIHW already: never scale `acc` in place
            if not self.bn_train:
                # folded eval-mode BatchNorm: d(gamma) = invstd * (<W, dW_folded> - mean * d(beta)), then dW = scale * dW_folded
                ops.bn_fold_grads(conv.w_src, w_oihw, a["dbeta"], conv.bn["running_mean"], conv.bn["running_var"],
                                  a["dgamma"], eps=conv.bn["eps"])
                ops.scale_rows(w_oihw, conv.bn_scale)
            grads[conv_name + ".weight"] = w_oihw
            grads[bn_name + ".weight"], grads[bn_name + ".bias"] = a["dgamma"], a["dbeta"]

        put(self.prefix + "0", self.prefix + "1", self.stem, acc["stem"], acc["stem"]["dw"])
        for p, (conv1, conv2, down), (a1, a2, ad) in zip(self.block_names, self.blocks, acc["blocks"]):
            put(p + "conv1", p + "bn1", conv1, a1, a1["dw"].permute(0, 3, 1, 2))
            put(p + "conv2", p + "bn2", conv2, a2, a2["dw"].permute(0, 3, 1, 2))
            if down is not None:
                put(p + "downsample.0", p + "downsample.1", down, ad, ad["dw"][:, 1:2, 1:2, :].permute(0, 3, 1, 2))


def build_backbone_plan(name, sd, bn_train=False):
    if bn_train and name != "res18":
        raise NotImplementedError(f"BatchNorm on batch statistics is implemented for ResNet-18 only, not {name!r}")
    with _PackBatch():                     # all of the backbone's filters packed by one launch
        if name == "vgg16":
            return VGG16Plan(sd)
        if name == "res18":
            return Res18Plan(sd, bn_train=bn_train)
        if name == "inv3":
            from .inception import Inv3Plan
            return Inv3Plan(sd)
    raise ValueError(f"backbone {name!r} is outside the hot-path scope (vgg16 / res18 / inv3)")


# ------------------------------------------------------------------------------------------------
# Dynamic inference module weights
# ------------------------------------------------------------------------------------------------
class DPIWeights:
    """One Dynamic_Person_Inference (dynamic_infer_module.py:14-151) in kernel layout."""

    def __init__(self, sd, prefix, kernel, ratios, scale_factor, beta_factor):
        self.kernel, self.ratios = tuple(kernel), list(ratios)
        self.scale_factor, self.beta_factor = bool(scale_factor), bool(beta_factor)
        self.taps = {}
        for r in self.ratios:
            self.taps[r] = ops.pack_din_weights(
                sd[f"{prefix}p_conv.{r}.weight"], sd[f"{prefix}p_conv.{r}.bias"],
                sd[f"{prefix}scale_conv.{r}.weight"] if scale_factor else None,
                sd[f"{prefix}scale_conv.{r}.bias"] if scale_factor else None)
        self.hidden = sd[f"{prefix}hidden_weight.weight"].contiguous().float()
        self.beta = sd[f"{prefix}beta"].contiguous().float() if beta_factor else None

    def __call__(self, x, out=None, accumulate=False, n_valid=None, tmp=None):
        """x [B,T,N,C] -> hidden_weight( mean_r / sum_r beta_r * DIN_r(x) );  out (+)= result."""
        if tmp is None:
            tmp = torch.zeros_like(x) if n_valid is not None else torch.empty_like(x)
        for i, r in enumerate(self.ratios):
            w_tap, b_cat = self.taps[r]
            if self.beta_factor:
                ops.dynamic_infer(x, w_tap, b_cat, self.kernel, r, scale_factor=self.scale_factor, out=tmp,
                                  coef_ptr=self.beta.data_ptr() + 4 * i, accumulate=i > 0, n_valid=n_valid)
            else:
                ops.dynamic_infer(x, w_tap, b_cat, self.kernel, r, scale_factor=self.scale_factor, out=tmp,
                                  coef=1.0 / len(self.ratios), accumulate=i > 0, n_valid=n_valid)
        return ops.linear_f32(tmp, self.hidden, None, out=out, accumulate=accumulate)


TCE_HEADS, TCE_DIM = 4, 128        # num_heads_context, num_features_context (infer_model.py:244-245)


def _context_position_table(oh, ow, device, downscale=16.0, num_pos_feats=256, temperature=10000.0):
    """Context_PositionEmbeddingSine(16, 512 / 2) (positional_encoding.py:67-93) as a [oh*ow, 512] table: a constant of the
    map size, evaluated once per plan (fp32, same op order as the reference)."""
    y = (torch.arange(1, oh + 1, dtype=torch.float32, device=device) * downscale).view(oh, 1).expand(oh, ow)
    x = (torch.arange(1, ow + 1, dtype=torch.float32, device=device) * downscale).view(1, ow).expand(oh, ow)
    j = torch.arange(num_pos_feats, dtype=torch.float32, device=device)
    dim_t = temperature ** (2 * torch.div(j, 2, rounding_mode="floor") / num_pos_feats)
    px, py = x[:, :, None] / dim_t, y[:, :, None] / dim_t
    px = torch.stack((px[:, :, 0::2].sin(), px[:, :, 1::2].cos()), dim=3).flatten(2)
    py = torch.stack((py[:, :, 0::2].sin(), py[:, :, 1::2].cos()), dim=3).flatten(2)
    return torch.cat((py, px), dim=2).reshape(oh * ow, 2 * num_pos_feats).contiguous()


class TCEWeights:
    """The context encoding of Dynamic_TCE_volleyball (one layer, four heads; TCE_STBiP_module.py:224-310) in kernel
    layout.  The heads' downsample2 1x1 convolutions are one tcgen05 GEMM over the feature map (N = 4 x 128, fp32 out);
    the position embedding enters as a per-pixel additive term posbias = downsample2(pos) + bias, computed once here
    (plan build: one-time weight preparation, like BatchNorm folding)."""

    def __init__(self, sd, prefix, out_size, device):
        H = TCE_HEADS
        w_all = torch.cat([sd[f"{prefix}{h}.downsample2.weight"] for h in range(H)], 0)            # [512, 512, 1, 1]
        b_all = torch.cat([sd[f"{prefix}{h}.downsample2.bias"] for h in range(H)], 0)
        if w_all.shape[1] != 512:
            raise ValueError("Dynamic_TCE_volleyball needs a 512-channel feature map (vgg16 / res18)")
        self.conv = _Conv(w_all.contiguous(), None, relu=False)
        self.w_all = w_all.view(H * TCE_DIM, 512).contiguous()                                      # fp32 [out, in]: backward
        oh, ow = out_size
        pos = _context_position_table(oh, ow, device)
        self.pos = pos                                                                              # [oh*ow, 512]: backward
        self.posbias = torch.addmm(b_all, pos, w_all.view(H * TCE_DIM, 512).t()).contiguous()       # [oh*ow, 512]
        g = lambda h, n: sd[f"{prefix}{h}.{n}"].contiguous()                                        # noqa: E731
        self.q = [(g(h, "emb_roi.weight"), g(h, "emb_roi.bias")) for h in range(H)]
        self.ln1 = [(g(h, "layernorm1.weight"), g(h, "layernorm1.bias")) for h in range(H)]
        self.ffn0 = [(g(h, "FFN.0.weight"), g(h, "FFN.0.bias")) for h in range(H)]
        self.ffn3 = [(g(h, "FFN.3.weight"), g(h, "FFN.3.bias")) for h in range(H)]
        self.ln2 = [(g(h, "layernorm2.weight"), g(h, "layernorm2.bias")) for h in range(H)]

    def __call__(self, x, fm, n):
        """x [B,T,N,NFB] fp32 person features, fm [F,OH,OW,512] fp16 feature map -> [B,T,N,NFB + 512]."""
        B, T, N, C = x.shape
        M, H, D = B * T * N, TCE_HEADS, TCE_DIM
        F_, oh, ow, _ = fm.shape
        if self.posbias.shape[0] != oh * ow:
            raise ValueError(f"feature map {oh}x{ow} does not match cfg.out_size (position table {self.posbias.shape[0]})")
        xf = x.reshape(M, C)
        q = torch.empty((H, M, D), dtype=torch.float32, device=x.device)
        for h in range(H):
            ops.linear_f32(xf, *self.q[h], out=q[h])                                    # emb_roi (:266)
        img = self.conv(fm, out_f32=True).view(F_, oh * ow, H * D)                       # downsample2 (:265), no bias
        ctx = ops.context_attention(q, img, self.posbias, N)                            # :274-280
        out = torch.empty((B, T, N, C + H * D), dtype=torch.float32, device=x.device)
        heads = []
        for h in range(H):
            c1 = ops.group_layernorm(ctx[h], *self.ln1[h], n_outer=M, outer_stride=D, cols=D, pre=q[h])     # :283
            f = ops.linear_f32(ops.linear_f32(c1, *self.ffn0[h], relu=True), *self.ffn3[h])                 # FFN :284
            heads.append(ops.group_layernorm(f, *self.ln2[h], n_outer=M, outer_stride=D, cols=D, pre=c1))   # :285
        # layout only: [person features | head 0 | ... | head 3] (torch.cat at :308 and infer_model.py:419)
        out.view(M, C + H * D)[:, :C].copy_(xf)
        for h in range(H):
            out.view(M, C + H * D)[:, C + h * D:C + (h + 1) * D].copy_(heads[h])
        return out


# ------------------------------------------------------------------------------------------------
# the whole path
# ------------------------------------------------------------------------------------------------
class DinEngine:
    """Forward plan for Dynamic_volleyball / Dynamic_collective built from a reference-named state_dict."""

    def __init__(self, cfg, state_dict, device, dataset="volleyball", frames_per_chunk=None, backbone_plan=None,
                 bn_train=False, tce=False):
        """backbone_plan: a plan built earlier from the same backbone weights (a training loop with the backbone
        frozen changes only the head's weights between steps: the 14.7 M backbone weights are not re-packed)."""
        self.cfg, self.dataset, self.device = cfg, dataset, torch.device(device)
        # frames per backbone launch: bounds the activation workspace (VGG-16 at 720p: 0.27 GB per frame
        # live at once) while keeping every launch many waves long
        self.frames_per_chunk = frames_per_chunk or int(os.environ.get("DIN_FRAMES_PER_CHUNK", "0")) or None
        sd = {k: _plan_tensor(v, self.device) for k, v in state_dict.items()}
        self.T, self.N = cfg.num_frames, cfg.num_boxes
        self.D, self.K = cfg.emb_features, cfg.crop_size[0]
        self.NFB = cfg.num_features_boxes
        self.C = cfg.lite_dim if cfg.lite_dim else self.NFB
        self.C_embed = self.C         # width of the person features embed() returns
        self.tce = None
        if tce:                       # Dynamic_TCE_volleyball: 4 x 128 context features per actor join the DIN input
            self.C = self.C + TCE_HEADS * TCE_DIM
        self.backbone_name = cfg.backbone
        self.backbone = (backbone_plan if backbone_plan is not None
                         else build_backbone_plan(cfg.backbone, sd, bn_train=bn_train))

        # fc_emb_1: the reference flattens crops as (d, ky, kx) (infer_model.py:181); RoIAlign here emits
        # (ky, kx, d), so permute the weight's columns once.  Public parameter stays [NFB, K*K*D].
        # The feature map's channel stride may exceed D (Inception: 1056 -> 1088, pad channels are zero), and
        # RoIAlign crops that stride, so the weight gets matching zero columns.
        self.D_stride = (self.D + 63) // 64 * 64
        w = sd["fc_emb_1.weight"].view(self.NFB, self.D, self.K * self.K).permute(0, 2, 1)
        wp = torch.zeros((self.NFB, self.K * self.K, self.D_stride), dtype=torch.float32, device=self.device)
        wp[:, :, :self.D] = w
        self.fc_emb = _Conv(wp.view(self.NFB, self.K * self.K * self.D_stride, 1, 1), sd["fc_emb_1.bias"], relu=False,
                            split=2 if cfg.backbone == "inv3" else 1)
        self.fc_emb_wk = wp.view(self.NFB, self.K * self.K * self.D_stride)   # fp32, kernel K order (d(crops) GEMM)
        self.fc_emb_bias = sd["fc_emb_1.bias"].contiguous().float()
        self.nl_emb = (sd["nl_emb_1.weight"].contiguous(), sd["nl_emb_1.bias"].contiguous())
        if cfg.lite_dim:
            pw = sd["point_conv.weight"]
            self.point_w = pw.reshape(pw.shape[0], pw.shape[1]).contiguous()
            self.point_b = sd["point_conv.bias"].contiguous()
            self.point_ln = (sd["point_ln.weight"].contiguous(), sd["point_ln.bias"].contiguous())
        self.hier = bool(getattr(cfg, "hierarchical_inference", False)) and dataset == "volleyball"
        sf, bf, ratios = cfg.scale_factor, cfg.beta_factor, cfg.sampling_ratio
        if dataset == "collective":
            self.dpis = [DPIWeights(sd, "DPI.", tuple(cfg.ST_kernel_size), ratios, sf, bf)]
        elif self.hier:
            k1, k2 = cfg.ST_kernel_size
            self.dpis = [DPIWeights(sd, "DPI.DPI_1.", k1, ratios, sf, bf),
                         DPIWeights(sd, "DPI.DPI_2.", k2, ratios, sf, bf)]
            self.hier_ln = (sd["DPI.hier_LN.weight"].contiguous(), sd["DPI.hier_LN.bias"].contiguous())
        else:
            self.dpis = [DPIWeights(sd, f"DPI.DIMlist.{i}.", cfg.ST_kernel_size[i], ratios, sf, bf)
                         for i in range(cfg.num_DIM)]
        self.dpi_nl = (sd["dpi_nl.weight"].contiguous(), sd["dpi_nl.bias"].contiguous())
        self.fc_act = (sd["fc_activities.weight"].contiguous(), sd["fc_activities.bias"].contiguous())
        self._idx_cache = {}
        self._fm_cache = None
        self._stage = None              # copy stream + rotating device buffers for frames handed over in host memory
        self._pending_small = None      # event of the boxes / actor counts staged on the copy stream (stage_small)
        if tce:
            self.tce = TCEWeights(sd, "multilayer_head_embfeature_context_encoding.CET.", tuple(cfg.out_size), self.device)

    # -- helpers ---------------------------------------------------------------------------------
    def _box_idx(self, n_frames, n_boxes):
        key = (n_frames, n_boxes)
        if key not in self._idx_cache:                # infer_model.py:155-157, built once per shape
            self._idx_cache[key] = torch.arange(n_frames, dtype=torch.int32, device=self.device) \
                .repeat_interleave(n_boxes).contiguous()
        return self._idx_cache[key]

    @staticmethod
    def _flat_frames(images):
        """[B,T,3,H,W] fp32 (volleyball.py:243,270) or [B,T,H,W,3] uint8 (the frame before the loader's
        transpose / float conversion) -> the same with B and T merged."""
        if images.dtype == torch.uint8:
            if images.dim() != 5 or images.shape[-1] != 3:
                raise ValueError(f"uint8 images must be [B,T,H,W,3], got {tuple(images.shape)}")
        elif images.dtype != torch.float32 or images.dim() != 5 or images.shape[2] != 3:
            raise ValueError(f"images must be fp32 [B,T,3,H,W] or uint8 [B,T,H,W,3], got {images.dtype} "
                             f"{tuple(images.shape)}")
        return images.reshape((-1,) + tuple(images.shape[2:])).contiguous()

    def features(self, images_flat):
        """[F,3,H,W] fp32 raw (or [F,H,W,3] uint8) -> NHWC fp16 [F,OH,OW,D] (prep_images + backbone), chunked
        over frames."""
        if getattr(self.backbone, "bn_train", False):        # BatchNorm on batch statistics: the training-mode forward
            return self.features_train(images_flat)[0]
        F_ = images_flat.shape[0]
        H, W = images_flat.shape[1:3] if images_flat.dtype == torch.uint8 else images_flat.shape[2:4]
        oh, ow, d = self.backbone.out_shape(H, W)
        assert d == self.D_stride, (d, self.D_stride)
        key = (F_, oh, ow, d)
        if self._fm_cache is None or self._fm_cache[0] != key:
            # zero-initialised once: pad channels beyond D are never written and must stay finite (zero)
            self._fm_cache = (key, torch.zeros(key, dtype=torch.float16, device=self.device))
        fm = self._fm_cache[1]
        per_chunk = frames_per_chunk(self.backbone_name, F_, H, W, self.frames_per_chunk)
        host = not images_flat.is_cuda
        _INFERENCE.on = True
        try:
            for f0 in range(0, F_, per_chunk):
                f1 = min(F_, f0 + per_chunk)
                chunk, slot = images_flat[f0:f1], None
                if host:
                    chunk, slot = self._stage_host_chunk(chunk, per_chunk)
                self.backbone(chunk, out=fm[f0:f1])                   # last layer writes its slice in place
                if slot is not None:
                    slot["freed"] = torch.cuda.Event()
                    slot["freed"].record(torch.cuda.current_stream())
        finally:
            _INFERENCE.on = False
        if self._pending_small is not None:                          # boxes / actor counts staged by stage_small()
            torch.cuda.current_stream().wait_event(self._pending_small)
            self._pending_small = None
        return fm

    def _stage_state(self):
        if self._stage is None:
            n_slots = max(2, int(os.environ.get("DIN_STAGE_SLOTS", "3")))
            self._stage = {"stream": torch.cuda.Stream(device=self.device), "slots": [{} for _ in range(n_slots)], "next": 0}
        return self._stage

    def stage_small(self, *tensors):
        """Small HOST tensors of a call (boxes, actor counts) -> device, on the copy stream and BEFORE the call's frames; the
        compute stream waits for them only after the backbone (end of features()).  Issued on the compute stream instead, a
        1 kB H2D copy queues behind the frame transfers in the copy engine and stalls the compute stream for up to a whole
        chunk transfer: measured 3.3 ms of idle GPU per 36 ms step (tools/probes/e2e_probe.py)."""
        st = self._stage_state()
        main = torch.cuda.current_stream()
        outs = []
        with torch.cuda.stream(st["stream"]):
            for t in tensors:
                o = t.to(self.device, non_blocking=True)
                o.record_stream(main)                                # allocated on the copy stream, consumed on `main`
                outs.append(o)
            ev = torch.cuda.Event()
            ev.record(st["stream"])
        self._pending_small = ev
        return outs

    def _stage_host_chunk(self, chunk, per_chunk):
        """Frames in (pinned) HOST memory: the chunk is copied to one of a few rotating device buffers on a copy stream and
        the compute stream waits for just that copy -- so the H2D transfer of chunk c + 1 (and, because nothing here
        blocks the host, of the next call's first chunks) runs under the backbone kernels of chunk c.  A step's inputs
        (885 MB of fp32 frames for 8 clips at 720p) then cost one chunk's transfer of pipeline fill, not the whole copy."""
        st = self._stage_state()
        slot = st["slots"][st["next"] % len(st["slots"])]
        st["next"] += 1
        main = torch.cuda.current_stream()
        shape = (per_chunk,) + tuple(chunk.shape[1:])
        buf = slot.get("buf")
        if buf is None or tuple(buf.shape) != shape or buf.dtype != chunk.dtype:
            buf = slot["buf"] = torch.empty(shape, dtype=chunk.dtype, device=self.device)
            # the block may be recycled memory that queued kernels of the compute stream still read: the copy stream
            # starts writing only after the compute stream has reached this point
            ev = torch.cuda.Event()
            ev.record(main)
            st["stream"].wait_event(ev)
            slot["freed"] = None
        dst = buf[:chunk.shape[0]]
        with torch.cuda.stream(st["stream"]):
            if slot.get("freed") is not None:
                st["stream"].wait_event(slot["freed"])            # the buffer's previous consumer has finished
            dst.copy_(chunk, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(st["stream"])
        main.wait_event(ready)
        return dst, slot

    def features_train(self, images_flat):
        """features() through the backbone plan's forward_train: -> (fm, [(f0, f1, saved)] per frame chunk)."""
        if not hasattr(self.backbone, "forward_train"):
            raise NotImplementedError(f"training the {self.backbone_name} backbone is not implemented (VGG-16 / ResNet-18)")
        F_ = images_flat.shape[0]
        H, W = images_flat.shape[1:3] if images_flat.dtype == torch.uint8 else images_flat.shape[2:4]
        oh, ow, d = self.backbone.out_shape(H, W)
        fm = torch.zeros((F_, oh, ow, d), dtype=torch.float16, device=images_flat.device)
        chunks = []
        # batch statistics couple all frames of the step: one chunk
        # training keeps every activation of a chunk for the backward: 16-frame chunks (VGG-16 at 720p: 0.56 GB / frame)
        per_chunk = F_ if getattr(self.backbone, "bn_train", False) else (self.frames_per_chunk or 16)
        for f0 in range(0, F_, per_chunk):
            f1 = min(F_, f0 + per_chunk)
            _, saved = self.backbone.forward_train(images_flat[f0:f1], out=fm[f0:f1])
            chunks.append((f0, f1, saved))
        return fm, chunks

    def embed(self, fm, boxes_flat, B, T, N):
        """RoIAlign -> fc_emb_1 -> nl_emb_1 -> ReLU -> (lite branch).  Returns fp32 [B,T,N,C]."""
        M = B * T * N
        if M < SMALL_EMBED_ROWS:
            # a handful of actors: the GEMM is a latency-bound sliver of one MMA tile either way, so it runs in fp32
            # on un-rounded crops and the fp32 weight (no fp16 rounding of either operand)
            crops = ops.roi_align_nhwc(fm, boxes_flat, self._box_idx(B * T, N), self.K, self.K, d=self.D_stride,
                                       out_f32=True)
            emb = ops.linear_f32(crops.view(M, -1), self.fc_emb_wk, self.fc_emb_bias)
        else:
            crops = ops.roi_align_nhwc(fm, boxes_flat, self._box_idx(B * T, N), self.K, self.K, d=self.D_stride)
            emb = self.fc_emb(crops.view(1, 1, M, self.K * self.K * self.D_stride), out_f32=True).view(M, self.NFB)
        x = ops.group_layernorm(emb, *self.nl_emb, n_outer=M, outer_stride=self.NFB, cols=self.NFB, relu=True)
        if self.cfg.lite_dim:
            y = ops.linear_f32(x, self.point_w, self.point_b)
            g = T * N * self.C_embed
            x = ops.group_layernorm(y, *self.point_ln, n_outer=B, outer_stride=g, cols=g, relu=True)
        return x.view(B, T, N, self.C_embed)

    # -- forwards --------------------------------------------------------------------------------
    @torch.no_grad()
    def forward_volleyball(self, images, boxes):
        B, T = images.shape[:2]
        N = self.N
        fm = self.features(self._flat_frames(images))
        OH, OW = self.cfg.out_size
        assert fm.shape[1:3] == (OH, OW), (tuple(fm.shape), self.cfg.out_size)          # infer_model.py:165
        x = self.embed(fm, boxes.reshape(B * T * N, 4).contiguous().float(), B, T, N)
        if self.tce is not None:
            x = self.tce(x, fm, N)                    # [B,T,N,C_person] -> [B,T,N,C_person + 512]  (infer_model.py:413-419)
        g_sz = T * N * self.C
        if self.hier:
            y1 = self.dpis[0](x)
            y1 = ops.group_layernorm(y1, *self.hier_ln, n_outer=B, outer_stride=g_sz, cols=g_sz, relu=True)
            g = self.dpis[1](y1)
        else:
            g = None
            for i, dpi in enumerate(self.dpis):
                g = dpi(x, out=g, accumulate=i > 0)
        if self.backbone_name == "res18":                                               # :203-209
            s = ops.group_layernorm(g, *self.dpi_nl, n_outer=B, outer_stride=g_sz, cols=g_sz, relu=True, post=x)
        else:                                                                           # :210-216
            s = ops.group_layernorm(g, *self.dpi_nl, n_outer=B, outer_stride=g_sz, cols=g_sz, relu=True, pre=x)
        return ops.readout(s, *self.fc_act)

    @torch.no_grad()
    def forward_collective(self, images, boxes, bboxes_num):
        B, T = images.shape[:2]
        N = self.N                                                                      # MAX_N
        fm = self.features(self._flat_frames(images))
        x = self.embed(fm, boxes.reshape(B * T * N, 4).contiguous().float(), B, T, N)
        n_valid = bboxes_num.reshape(B, T)[:, 0].to(torch.int32).contiguous()           # :1289
        g = self.dpis[0](x, n_valid=n_valid)
        # (g + x) -> [N, T, C] -> LayerNorm([T, C]) -> ReLU  (:1298-1301): one group per (clip, actor)
        s = torch.zeros_like(g)
        ops.group_layernorm(g, *self.dpi_nl, n_outer=B, n_inner=N, outer_stride=T * N * self.C,
                            inner_stride=self.C, rows=T, row_stride=N * self.C, cols=self.C, relu=True, pre=x,
                            n_valid=n_valid, out=s)
        return ops.readout(s, *self.fc_act, n_valid=n_valid)


class BasenetEngine:
    """Forward plan of the stage-1 base models (reference base_model.py:64-142, 200-284): the stage-2 path's
    backbone + RoIAlign + embedding GEMM (ReLU fused in its epilogue), then fc_actions per actor and
    max-over-actors -> fc_activities."""

    def __init__(self, cfg, state_dict, device, dataset="volleyball", emb_name="fc_emb"):
        self.cfg, self.dataset, self.device = cfg, dataset, torch.device(device)
        self.frames_per_chunk = int(os.environ.get("DIN_FRAMES_PER_CHUNK", "0")) or None
        sd = {k: _plan_tensor(v, self.device) for k, v in state_dict.items()}
        self.N, self.D, self.K, self.NFB = cfg.num_boxes, cfg.emb_features, cfg.crop_size[0], cfg.num_features_boxes
        self.backbone_name = "inv3" if dataset == "collective" else cfg.backbone       # base_model.py:159
        self.backbone = build_backbone_plan(self.backbone_name, sd)
        self.D_stride = (self.D + 63) // 64 * 64
        w = sd[emb_name + ".weight"].view(self.NFB, self.D, self.K * self.K).permute(0, 2, 1)
        wp = torch.zeros((self.NFB, self.K * self.K, self.D_stride), dtype=torch.float32, device=self.device)
        wp[:, :, :self.D] = w
        self.fc_emb = _Conv(wp.view(self.NFB, self.K * self.K * self.D_stride, 1, 1), sd[emb_name + ".bias"],
                            relu=True, split=2 if self.backbone_name == "inv3" else 1)
        self.fc_emb_wk = wp.view(self.NFB, self.K * self.K * self.D_stride)   # fp32, kernel K order (d(crops) GEMM)
        self.emb_name = emb_name
        self.fc_actions = (sd["fc_actions.weight"].contiguous(), sd["fc_actions.bias"].contiguous())
        self.fc_act = (sd["fc_activities.weight"].contiguous(), sd["fc_activities.bias"].contiguous())
        self._idx_cache, self._fm_cache = {}, None
        self._stage, self._pending_small = None, None

    _box_idx = DinEngine._box_idx
    _flat_frames = staticmethod(DinEngine._flat_frames)
    features = DinEngine.features
    features_train = DinEngine.features_train
    _stage_state = DinEngine._stage_state
    _stage_host_chunk = DinEngine._stage_host_chunk
    stage_small = DinEngine.stage_small

    def _states(self, images, boxes, B, T, N):
        fm = self.features(self._flat_frames(images))
        OH, OW = self.cfg.out_size
        assert fm.shape[1:3] == (OH, OW), (tuple(fm.shape), self.cfg.out_size)
        M = B * T * N
        crops = ops.roi_align_nhwc(fm, boxes.reshape(M, 4).contiguous().float(), self._box_idx(B * T, N), self.K,
                                   self.K, d=self.D_stride)
        # fc_emb + ReLU (base_model.py:119-120): ReLU fused in the GEMM epilogue, fp32 out
        return self.fc_emb(crops.view(1, 1, M, self.K * self.K * self.D_stride), out_f32=True).view(M, self.NFB)

    @torch.no_grad()
    def forward_volleyball(self, images, boxes):
        B, T = images.shape[:2]
        N = self.N
        x = self._states(images, boxes, B, T, N)                                        # [B*T*N, NFB]
        actions = ops.linear_f32(x, *self.fc_actions)                                   # :129
        activities = ops.readout(x.view(B, T, N, self.NFB), *self.fc_act)               # :133-136,140 (mean over T)
        if T != 1:                                                                      # :138-139
            actions = ops.mean_axis(actions.view(B, T, N * actions.shape[-1]), 1)
        return actions.view(B * N, -1), activities

    @torch.no_grad()
    def forward_collective(self, images, boxes, bboxes_num):
        B, T = images.shape[:2]
        N = self.N                                                                      # MAX_N
        x = self._states(images, boxes, B, T, N)
        n_valid = bboxes_num.reshape(B * T).to(torch.int32).contiguous()                # per FRAME (:252)
        actions_all = ops.linear_f32(x, *self.fc_actions)                               # [B*T*MAX_N, A]
        activities = ops.readout(x.view(B * T, 1, N, self.NFB), *self.fc_act, n_valid=n_valid)   # [B*T, A]
        # the reference concatenates the first N_bt rows of every frame (:256-279): a ragged gather whose
        # size the host must know, exactly as the reference's Python loop does (one sync on bboxes_num)
        keep = (torch.arange(N, device=self.device).view(1, N) < n_valid.view(B * T, 1)).reshape(-1)
        return actions_all[keep.nonzero(as_tuple=True)[0]], activities
