"""Drop-in for `backbone.backbone` (reference backbone/backbone.py:10-132).

The classes are PARAMETER CONTAINERS with the reference's state_dict layout (they hold the same
torchvision sub-modules, so `backbone_state_dict` checkpoints written by base_model.py:46-55 load
unchanged).  Their forward runs the sm_100a plan in din_b200/engine.py — tcgen05 implicit-GEMM
convolutions on NHWC fp16 — never the torchvision modules.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F          # noqa: F401  (the reference's star-imports hand `F` to its trainers)
import torchvision.models as models

from din_b200 import engine as _engine


def _build(fn, pretrained, **kw):
    """torchvision constructor.  `pretrained=True` means ImageNet weights exactly as in the reference
    (`models.vgg16(pretrained=True)`, backbone.py:14,92,118): torchvision loads them from the hub cache or downloads
    them, and a failed download RAISES -- stage-1 training silently starting from random weights would be far worse
    than an error.  Machines without network access opt out explicitly with DIN_OFFLINE=1 (tests, the synthetic
    benchmark, stage-2 runs that overwrite the backbone through loadmodel() anyway): then the cached checkpoint is
    used when present and random initialisation otherwise."""
    if not pretrained:
        return fn(weights=None, **kw)
    w = models.get_model_weights(fn).DEFAULT
    if os.environ.get("DIN_OFFLINE", "0") not in ("", "0"):
        path = os.path.join(torch.hub.get_dir(), "checkpoints", os.path.basename(w.url))
        if not os.path.exists(path):
            return fn(weights=None, **kw)
    return fn(weights=w, **kw)


class _PlanBackbone(nn.Module):
    _plan_name = None

    def _plan(self):
        key = tuple((p.data_ptr(), p._version) for p in self.state_dict().values())
        if getattr(self, "_plan_key", None) != key:
            sd = {"backbone." + k: v.detach().float() if v.is_floating_point() else v
                  for k, v in self.state_dict().items()}
            self._plan_obj = _engine.build_backbone_plan(self._plan_name, sd)
            self._plan_key = key
        return self._plan_obj

    def forward_nhwc(self, x_raw):
        """raw fp32 NCHW images (0..255) -> NHWC fp16 feature map(s); prep_images is fused in."""
        if not x_raw.is_cuda:
            raise RuntimeError("backbone: CUDA tensors required (no CPU fallback on the DIN hot path)")
        return self._plan()(x_raw.contiguous())


class MyVGG16(_PlanBackbone):
    _plan_name = "vgg16"

    def __init__(self, pretrained=False):
        super().__init__()
        self.features = _build(models.vgg16, pretrained).features

    def forward(self, x):
        """x: images AFTER prep_images (reference contract) -> [NCHW fp32 feature map]."""
        raw = ((x / 2.0) + 0.5) * 255.0
        return [self.forward_nhwc(raw).permute(0, 3, 1, 2).float()]


class MyRes18(_PlanBackbone):
    _plan_name = "res18"

    def __init__(self, pretrained=False):
        super().__init__()
        r = _build(models.resnet18, pretrained)
        self.features = nn.Sequential(r.conv1, r.bn1, r.relu, r.maxpool, r.layer1, r.layer2, r.layer3, r.layer4)

    def forward(self, x):
        raw = ((x / 2.0) + 0.5) * 255.0
        return [self.forward_nhwc(raw).permute(0, 3, 1, 2).float()]


class MyInception_v3(_PlanBackbone):
    _plan_name = "inv3"
    _names = ["Conv2d_1a_3x3", "Conv2d_2a_3x3", "Conv2d_2b_3x3", "Conv2d_3b_1x1", "Conv2d_4a_3x3", "Mixed_5b",
              "Mixed_5c", "Mixed_5d", "Mixed_6a", "Mixed_6b", "Mixed_6c", "Mixed_6d", "Mixed_6e"]

    def __init__(self, transform_input=False, pretrained=False):
        super().__init__()
        if transform_input:
            raise NotImplementedError("transform_input=True is never used by the reference models")
        self.transform_input = transform_input
        inc = _build(models.inception_v3, pretrained, aux_logits=True, init_weights=False)
        for n in self._names:
            setattr(self, n, getattr(inc, n))

    def forward(self, x):
        raw = ((x / 2.0) + 0.5) * 255.0
        fm = self.forward_nhwc(raw)                       # multiscale map; channels [0, 288) = Mixed_5d
        out1 = self._plan().last_out1                     # Mixed_6e before the upsample
        return [fm[..., :288].permute(0, 3, 1, 2).float(), out1.permute(0, 3, 1, 2).float()]


def _out_of_scope(name):
    class _Stub(nn.Module):
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name} is outside the DIN stage-2 hot-path scope (SURVEY.md §2 #2)")
    _Stub.__name__ = name
    return _Stub


MyVGG19 = _out_of_scope("MyVGG19")
MyRes50 = _out_of_scope("MyRes50")
MyAlex = _out_of_scope("MyAlex")
