"""ref_harness.py — imports the REAL reference classes from /root/reference.  TEST INFRASTRUCTURE ONLY.

Usable where /root/reference exists (the authoring container) or where oracle/make_ref.py staged a byte-for-byte
copy under oracle/_ref/reference (git-ignored; it travels to the GPU box with the snapshot).  It is how the
oracle restatement (din_oracle.py) is pinned and how the fixtures in tests/golden/ are generated.

Recipe:
  * sys.path = [oracle/shims, /root/reference, ...] for the duration of the import, with the colliding
    top-level module names (config, utils, backbone, infer_model, infer_module, ...) isolated in
    sys.modules so that this repo's drop-in modules of the same names are never shadowed or polluted;
  * torchvision.models.{vgg16,...}(pretrained=True) is intercepted to weights=None (no network);
  * reference defects that make 3 of 5 BASELINE configs crash are patched minimally:
      patch I  (inv3): Dynamic_volleyball.forward has fusion branches only for res18 / vgg16
               (infer_model.py:203-216) -> the model is built with cfg.backbone='inv3' and cfg.backbone is
               then set to 'vgg16', which forward() reads only at the fusion branch;
      patch H  (hierarchical): Hierarchical_Dynamic_Inference.forward (dynamic_infer_module.py:491-498)
               feeds DPI_1's tuple to LayerNorm, returns one value where two are unpacked, and applies
               F.dropout in eval mode -> replaced by a forward that takes [0], honours self.training and
               returns (out, mad);
      patch C  (collective): Dynamic_collective.forward adds DPI's tuple to a tensor
               (infer_model.py:1294,1298) -> model.DPI is wrapped so it returns its first output.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = next((p for p in ("/root/reference", os.path.join(_HERE, "_ref", "reference"))
                 if os.path.exists(os.path.join(p, "infer_model.py"))), "/root/reference")
_SHIMS = os.path.join(_HERE, "shims")

_COLLIDING = ("config", "utils", "backbone", "infer_model", "infer_module", "gcn_model", "base_model",
              "roi_align", "thop", "fvcore", "skimage", "dataset", "volleyball", "collective",
              "train_net_dynamic", "train_net")

_ref_modules: dict = {}


def available() -> bool:
    return os.path.isdir(REF_ROOT) and os.path.exists(os.path.join(REF_ROOT, "infer_model.py"))


@contextlib.contextmanager
def _isolated_import():
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in _COLLIDING}
    for k in saved:
        del sys.modules[k]
    sys.modules.update(_ref_modules)
    saved_path = list(sys.path)
    # The reference's infer_module/ and backbone/ have no __init__.py (namespace packages); a regular
    # package of the same name anywhere on sys.path would win, so this repo's drop-in package directory
    # is taken off the path while reference modules are being imported.
    pkg = os.path.join(os.path.dirname(_HERE), "din-group-activity-recognition-benchmark_b200")
    sys.path[:] = [_SHIMS, REF_ROOT] + [p for p in sys.path if os.path.abspath(p or ".") != pkg]
    import torchvision.models as tvm
    orig = {}

    def _wrap(name):
        fn = getattr(tvm, name)
        orig[name] = fn

        def build(*a, pretrained=False, **k):
            k.setdefault("weights", None)
            if name == "inception_v3":
                k.setdefault("aux_logits", True)
                k.setdefault("init_weights", False)
            return fn(*a, **k)
        setattr(tvm, name, build)

    for n in ("vgg16", "vgg19", "resnet18", "resnet50", "alexnet", "inception_v3"):
        _wrap(n)
    try:
        yield
    finally:
        for n, fn in orig.items():
            setattr(tvm, n, fn)
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k.split(".")[0] in _COLLIDING]:
            _ref_modules[k] = sys.modules.pop(k)
        sys.modules.update(saved)


def ref_module(name: str) -> types.ModuleType:
    """Import a reference module by its flat name (e.g. 'infer_model') in isolation."""
    if not available():
        raise RuntimeError("/root/reference is not present on this machine")
    with _isolated_import():
        import importlib
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return importlib.import_module(name)


def make_ref_cfg(pc):
    """oracle PathConfig -> a reference Config object with the script's fields set
    (scripts/train_volleyball_stage2_dynamic.py:17-40, scripts/train_collective_stage2_dynamic.py)."""
    Config = ref_module("config").Config
    cfg = Config(pc.dataset)
    cfg.log_path = None            # utils.print_log then only prints (utils.py:101-105)
    cfg.backbone = pc.backbone
    cfg.image_size = tuple(pc.image_size)
    cfg.out_size = tuple(pc.out_size)
    cfg.emb_features = pc.emb_features
    cfg.num_frames = pc.num_frames
    cfg.num_boxes = pc.num_boxes
    cfg.crop_size = tuple(pc.crop_size)
    cfg.num_features_boxes = pc.num_features_boxes
    cfg.num_features_gcn = pc.num_features_boxes
    cfg.num_activities = pc.num_activities
    cfg.lite_dim = pc.lite_dim
    cfg.ST_kernel_size = pc.ST_kernel_size
    cfg.sampling_ratio = list(pc.sampling_ratio)
    cfg.scale_factor = pc.scale_factor
    cfg.beta_factor = pc.beta_factor
    cfg.hierarchical_inference = pc.hierarchical_inference
    cfg.num_DIM = pc.num_DIM
    cfg.dynamic_sampling = True
    cfg.parallel_inference = False
    cfg.stride = 1
    cfg.group = 1
    cfg.train_backbone = False
    cfg.training_stage = 2
    return cfg


def _patched_hier_forward(self, person_features):       # patch H
    p1 = self.DPI_1(person_features)[0]
    p1 = F.relu(self.hier_LN(p1))
    p1 = F.dropout(p1, training=self.training)
    return self.DPI_2(p1)


class _FirstOutput(nn.Module):                          # patch C
    def __init__(self, inner):
        super().__init__()
        self.inner = inner

    def forward(self, x):
        return self.inner(x)[0]


def build_ref_model(pc, sd):
    """Build the reference's Dynamic_volleyball / Dynamic_collective for PathConfig `pc`, load `sd`
    (reference key names), apply the documented patches, return it in eval mode."""
    im = ref_module("infer_model")
    cfg = make_ref_cfg(pc)
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        with _isolated_import():
            cls = im.Dynamic_collective if pc.dataset == "collective" else \
                (im.Dynamic_TCE_volleyball if getattr(pc, "tce", False) else im.Dynamic_volleyball)
            model = cls(cfg)
    missing, unexpected = model.load_state_dict(sd, strict=True), None
    model.eval()
    if pc.hierarchical_inference:
        model.DPI.forward = types.MethodType(_patched_hier_forward, model.DPI)
    if pc.dataset == "collective":
        model.DPI = _FirstOutput(model.DPI)
    if pc.backbone not in ("res18", "vgg16"):
        cfg.backbone = "vgg16"                           # patch I (read only at infer_model.py:203,210)
    return model


def ref_forward(pc, sd, *batch):
    model = build_ref_model(pc, sd)
    import io
    import warnings
    with torch.no_grad(), warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        return model(tuple(batch))["activities"]


def ref_head_grads(pc, sd, labels, *batch, train_backbone=False, bn_train=False, return_buffers=False):
    """Train-mode forward + F.cross_entropy + backward of the REFERENCE model (train_net_dynamic.py:170-224)
    with the backbone frozen (config.py:39) or trained (scripts/train_volleyball_stage2_dynamic.py:12), BatchNorm
    layers in eval mode (train_net_dynamic.py:101-102 `set_bn_eval`) -- or, with bn_train, on batch statistics
    (cfg.set_bn_eval = False, the config.py:80 default) -- and train_dropout_prob = 0 (deterministic).
    -> (logits, loss, {param name: grad}) [, {BatchNorm buffer name: value after the step}]."""
    model = build_ref_model(pc, sd)
    for q in model.backbone.parameters():
        q.requires_grad = bool(train_backbone)
    model.cfg.train_dropout_prob = 0.0
    model.train()
    model.dropout_global.p = 0.0
    for m in model.modules():
        if isinstance(m, nn.Dropout) or (isinstance(m, nn.BatchNorm2d) and not bn_train):
            m.eval()
    import io
    import warnings
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        logits = model(tuple(batch))["activities"]
        loss = F.cross_entropy(logits, labels)
        loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    grads = {(n.replace("DPI.inner.", "DPI.") if pc.dataset == "collective" else n): g for n, g in grads.items()}
    if return_buffers:
        bufs = {n: b.detach().clone() for n, b in model.named_buffers() if n.startswith("backbone.")}
        return logits.detach(), loss.detach(), grads, bufs
    return logits.detach(), loss.detach(), grads


def ref_basenet_forward(pc, sd, *batch):
    """The reference's stage-1 Basenet_volleyball / Basenet_collective (base_model.py) in eval mode.
    Basenet_collective hard-codes MyInception_v3 (base_model.py:159); no patches are needed."""
    bm = ref_module("base_model")
    cfg = make_ref_cfg(pc)
    cfg.num_actions = pc.num_actions
    import io
    import warnings
    with contextlib.redirect_stdout(io.StringIO()):
        with _isolated_import():
            model = (bm.Basenet_collective if pc.dataset == "collective" else bm.Basenet_volleyball)(cfg)
    model.load_state_dict(sd, strict=True)
    model.eval()
    with torch.no_grad(), warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        return model(tuple(batch))


def ref_basenet_grads(pc, sd, actions_labels, activities_labels, actions_weights, *batch):
    """The reference's stage-1 training step (train_net.py:163-189) on base_model.Basenet_volleyball: train mode,
    dropout 0, everything trainable.  -> (loss, {param name: grad})."""
    bm = ref_module("base_model")
    cfg = make_ref_cfg(pc)
    cfg.num_actions = pc.num_actions
    cfg.train_dropout_prob = 0.0
    cfg.train_backbone = True
    import io
    import warnings
    with contextlib.redirect_stdout(io.StringIO()):
        with _isolated_import():
            model = bm.Basenet_volleyball(cfg)
    model.load_state_dict(sd, strict=True)
    model.train()
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        actions, activities = model(tuple(batch))
        loss = F.cross_entropy(activities, activities_labels) + \
            cfg.actions_loss_weight * F.cross_entropy(actions, actions_labels, weight=actions_weights)
        loss.backward()
    return loss.detach(), {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}


def ref_dpi_module(in_dim, kernel, ratios, scale_factor=True, beta_factor=False):
    """A bare reference Dynamic_Person_Inference (dynamic_infer_module.py:14-404)."""
    dm = ref_module("infer_module.dynamic_infer_module")
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return dm.Dynamic_Person_Inference(in_dim=in_dim, person_mat_shape=(10, 12), stride=1,
                                           kernel_size=kernel, dynamic_sampling=True,
                                           sampling_ratio=list(ratios), group=1, scale_factor=scale_factor,
                                           beta_factor=beta_factor, parallel_inference=False, cfg=None).eval()
