"""Drop-in for the DIN models of `infer_model` (reference infer_model.py:15-234, 1135-1319).

`Dynamic_volleyball(cfg)` / `Dynamic_collective(cfg)` keep the reference's constructor contract, module
tree and state_dict key names/shapes, `loadmodel`, and `forward(batch) -> {'activities'}`,
so `scripts/train_volleyball_stage2_dynamic.py`'s model registry resolves and stage-1 / stage-2
checkpoints load unchanged.  forward() runs the sm_100a plan (din_b200/engine.py); torch modules below
are parameter containers and are never called.

Scope: evaluation / inference forward, and the training step: `model.train()` + `loss.backward()` produce
gradients for every parameter after the backbone through the backward kernels in csrc/head_bwd.cu, and -- with
cfg.train_backbone = True and the VGG-16, ResNet-18 or Inception-v3 backbone (BatchNorm in eval mode) -- for the backbone too
(RoIAlign scatter, ReLU / max-pool backward, dgrad on the forward tcgen05 kernel, wgrad on
csrc/conv_wgrad_tcgen05.cu, zero insertion for the stride-2 layers).
ResNet-18 also trains with BatchNorm on batch statistics (csrc/bn_train.cu: no cfg.set_bn_eval, the reference's default).
Inception-v3 with BatchNorm on batch statistics is not implemented and raises (freeze BN as cfg.set_bn_eval does).
"""
import collections

# The reference's trainers take `torch`, `nn`, `F`, `np`, `models`, the meters ... from the star-imports of this module
# (train_net_dynamic.py:13 `from infer_model import *`, infer_model.py:1-2), so the same two star-imports stay here.
from backbone.backbone import *          # noqa: F401,F403
from utils import *                      # noqa: F401,F403

import torch
import torch.nn as nn

from backbone.backbone import MyInception_v3, MyRes18, MyVGG16
from din_b200 import plan_cache as _pc
from din_b200 import train as _train
from din_b200.engine import DinEngine
from infer_module.dynamic_infer_module import (Dynamic_Person_Inference, Hierarchical_Dynamic_Inference,
                                               Multi_Dynamic_Inference)
from roi_align.roi_align import RoIAlign
from utils import print_log


def _as_frames(images):
    """fp32 [B,T,3,H,W] as the reference loader yields it (other float dtypes are converted); uint8
    [B,T,H,W,3] (decoded frames, no transpose / float conversion) goes to the uint8-ingest stem as is."""
    return images if images.dtype == torch.uint8 else images.float()


def _make_backbone(cfg):
    if cfg.backbone == "inv3":
        return MyInception_v3(transform_input=False, pretrained=True)
    if cfg.backbone == "vgg16":
        return MyVGG16(pretrained=True)
    if cfg.backbone == "res18":
        return MyRes18(pretrained=True)
    raise NotImplementedError(f"backbone {cfg.backbone!r} is outside the DIN hot-path scope")


class _DinTrainFn(torch.autograd.Function):
    """The whole CUDA path as ONE autograd node: forward = din_b200.train.forward_train, backward =
    din_b200.train.backward_head (gradients under the reference's parameter names)."""

    @staticmethod
    def forward(ctx, model, images, boxes, bboxes_num, names, *params):
        logits, tape = _train.forward_train(model.engine(), images, boxes, bboxes_num, training=model.training,
                                            train_backbone=any(n.startswith("backbone.") for n in names))
        ctx.eng, ctx.tape, ctx.names = model.engine(), tape, names
        ctx.sink = getattr(_pc.owner_of(model), "grad_sink", None)
        model._last_tape = tape if getattr(model, "keep_tape", False) else None     # test / debugging hook
        ctx.shapes = [tuple(p.shape) for p in params]
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        sink = ctx.sink if (ctx.sink is not None and ctx.sink.active()) else None
        grads = _train.backward_head(ctx.eng, ctx.tape, dlogits, sink=sink)
        ctx.tape = None
        if sink is not None:
            # data-parallel run (din_b200.parallel.BucketedGradientReducer): the gradients were packed into the flat
            # buffer and all-reduced as the backward produced them; .grad becomes a view of that buffer (no copies)
            sink.finish()
            return (None,) * (5 + len(ctx.names))
        out = [grads[n].reshape(shp) if (need and n in grads) else None
               for n, shp, need in zip(ctx.names, ctx.shapes, ctx.needs_input_grad[5:])]
        return (None, None, None, None, None) + tuple(out)


class _DinModel(nn.Module):
    _dataset = None
    _tce = False

    def _common_init(self, cfg, person_mat_shape, ln_shape):
        self.cfg = cfg
        T, N = cfg.num_frames, cfg.num_boxes
        D, K, NFB = cfg.emb_features, cfg.crop_size[0], cfg.num_features_boxes
        self.backbone = _make_backbone(cfg)
        if not cfg.train_backbone:
            for p in self.backbone.parameters():
                p.requires_grad = False
        self.roi_align = RoIAlign(*cfg.crop_size)
        self.fc_emb_1 = nn.Linear(K * K * D, NFB)
        self.nl_emb_1 = nn.LayerNorm([NFB])
        in_dim = cfg.lite_dim if cfg.lite_dim else NFB
        print_log(cfg.log_path, ("Activate" if cfg.lite_dim else "Deactivate") + " lite model inference.")
        person_dim = in_dim
        if self._tce:
            # TCE module (infer_model.py:283-289): 4 heads x 128 context features join the DIN input (:292)
            if cfg.lite_dim or N != 12 or cfg.backbone not in ("vgg16", "res18"):
                raise NotImplementedError(
                    "Dynamic_TCE_volleyball: the reference's context encoder takes NFB-dimensional person features "
                    "(lite_dim must be None), asserts 12 actors (TCE_STBiP_module.py:261) and a 512-channel feature map "
                    "(vgg16 / res18)")
            self.multilayer_head_embfeature_context_encoding = _ContextEncoding(4, 128, NFB)
            in_dim = in_dim + 4 * 128
        kw = dict(in_dim=in_dim, person_mat_shape=person_mat_shape, stride=cfg.stride,
                  kernel_size=cfg.ST_kernel_size, dynamic_sampling=cfg.dynamic_sampling,
                  sampling_ratio=cfg.sampling_ratio, group=cfg.group, scale_factor=cfg.scale_factor,
                  beta_factor=cfg.beta_factor, parallel_inference=cfg.parallel_inference, cfg=cfg)
        if cfg.hierarchical_inference:
            self.DPI = Hierarchical_Dynamic_Inference(**kw)
        elif self._dataset == "volleyball":
            self.DPI = Multi_Dynamic_Inference(num_DIM=cfg.num_DIM, **kw)
        else:
            self.DPI = Dynamic_Person_Inference(**kw)
        print_log(cfg.log_path, "Hierarchical Inference : " + str(cfg.hierarchical_inference))
        self.dpi_nl = nn.LayerNorm(ln_shape(T, N, in_dim))
        self.dropout_global = nn.Dropout(p=cfg.train_dropout_prob)
        if cfg.lite_dim:
            self.point_conv = nn.Conv2d(NFB, person_dim, kernel_size=1, stride=1)
            self.point_ln = nn.LayerNorm([T, N, person_dim])
        self.fc_activities = nn.Linear(in_dim, cfg.num_activities)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.kaiming_normal_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        self._owner = [self]               # nn.DataParallel replicas find the original model here (din_b200/plan_cache.py)
        self._plans = _pc.PlanTable()

    # -- reference API ---------------------------------------------------------------------------
    def loadmodel(self, filepath):
        state = torch.load(filepath)
        self.backbone.load_state_dict(state["backbone_state_dict"])
        self.fc_emb_1.load_state_dict(state["fc_emb_state_dict"])
        print("Load model states from: ", filepath)

    def loadpart(self, pretrained_state_dict, model, prefix):
        own = model.state_dict()
        picked = collections.OrderedDict((k.replace(prefix, ""), v) for k, v in pretrained_state_dict.items()
                                         if k.replace(prefix, "") in own)
        own.update(picked)
        model.load_state_dict(own)
        print(str(len(picked)) + " parameters loaded for " + prefix)

    # -- plan management -------------------------------------------------------------------------
    def _bn_batch_stats(self):
        """True when the backbone's BatchNorm layers run on batch statistics (train() without cfg.set_bn_eval)."""
        return self.training and any(isinstance(m, nn.modules.batchnorm._BatchNorm) and m.training
                                     for m in self.backbone.modules())

    def engine(self):
        """The forward plan for this model on its device: rebuilt only when a tensor of the (owner) model changed."""
        own = _pc.owner_of(self)
        dev = _pc.device_of(self)
        bn_train = self._bn_batch_stats()
        key = (bn_train,) + _pc.version_key(self)
        slot = own._plans.get(dev)
        if slot is None or slot["key"] != key:
            _pc.require_cuda(dev)
            # the backbone plan survives head-only weight updates (frozen-backbone training); on batch statistics the
            # convolutions run un-folded and the kernels read the running statistics through live pointers, so the
            # BatchNorm buffers (bumped every step) are not part of the backbone's key
            bb_key = (bn_train,) + _pc.version_key(self, buffers=not bn_train, prefix="backbone")
            plan = slot["engine"].backbone if (slot is not None and slot["bb_key"] == bb_key) else None
            with torch.cuda.device(dev):
                eng = DinEngine(self.cfg, _pc.named_tensors(self), dev, dataset=self._dataset, backbone_plan=plan,
                                bn_train=bn_train, tce=self._tce)
            slot = own._plans.put(dev, key=key, bb_key=bb_key, engine=eng)
        return slot["engine"]

    def _check_mode(self, images):
        """-> the device the plan runs on.  Frames may also arrive in (pinned) HOST memory in eval mode: they are streamed to
        the GPU chunk by chunk under the backbone's kernels (DinEngine._stage_host_chunk) -- a transport, not a CPU
        fallback: the model itself must live on a CUDA device."""
        dev = self.fc_activities.weight.device
        if dev.type != "cuda":
            raise RuntimeError("the DIN hot path runs on sm_100a only: move the model to a CUDA device (there is no CPU "
                               "fallback)")
        if not images.is_cuda and self.training:
            raise RuntimeError("host-memory frames are streamed in eval mode only: pass CUDA tensors to a model in train()")
        return dev

    def _to_device(self, dev, *small):
        """Boxes / actor counts handed over in host memory: in eval mode they travel on the engine's copy stream ahead of
        the frames (DinEngine.stage_small); in train mode a plain copy."""
        if self.training:
            return [t.to(dev) for t in small]
        return self.engine().stage_small(*small)

    def _run(self, images, boxes, bboxes_num=None):
        """eval: the forward plan.  train: forward with dropout; with grad enabled additionally one autograd
        node whose backward is the CUDA head backward."""
        if self._bn_batch_stats():
            bns = [m for m in self.backbone.modules() if isinstance(m, nn.modules.batchnorm._BatchNorm)]
            if self.cfg.backbone != "res18" or not all(m.training and m.momentum == 0.1 and m.track_running_stats
                                                       for m in bns):
                raise NotImplementedError(
                    "train mode with BatchNorm batch statistics is implemented for the ResNet-18 backbone (all of its "
                    "BatchNorm layers in train mode): freeze BN as the reference does with cfg.set_bn_eval "
                    "(train_net_dynamic.py:101-102, model.apply(set_bn_eval))")
        eng = self.engine()
        if self._tce:                                       # context_dropout_ratio of the encoder (TCE_STBiP_module.py:225)
            eng.tce_dropout = float(self.multilayer_head_embfeature_context_encoding.CET[0].dropout.p)
        if not self.training:
            if self._dataset == "collective":
                return eng.forward_collective(images, boxes, bboxes_num)
            return eng.forward_volleyball(images, boxes)
        if not torch.is_grad_enabled():
            return _train.forward_train(eng, images, boxes, bboxes_num, training=True)[0]
        named = _pc.trainable(self)
        if any(n.startswith("backbone.") for n, _ in named) and self.cfg.backbone not in ("vgg16", "res18", "inv3"):
            raise NotImplementedError(
                f"training the backbone is implemented for VGG-16, ResNet-18 and Inception-v3 (BatchNorm in eval mode), "
                f"not {self.cfg.backbone!r}: set cfg.train_backbone = False (config.py:39, the stage-2 default) -- "
                "SURVEY.md §8f rank 1")
        names = tuple(n for n, _ in named)
        return _DinTrainFn.apply(self, images, boxes, bboxes_num, names, *[p for _, p in named])


class Dynamic_volleyball(_DinModel):
    """reference infer_model.py:15-234."""
    _dataset = "volleyball"

    def __init__(self, cfg):
        super().__init__()
        self._common_init(cfg, (10, 12), lambda T, N, C: [T, N, C])     # person_mat_shape hard-coded :77

    def forward(self, batch_data):
        images_in, boxes_in = batch_data
        dev = self._check_mode(images_in)
        with torch.cuda.device(dev):
            if not boxes_in.is_cuda:
                boxes_in, = self._to_device(dev, boxes_in)
            scores = self._run(_as_frames(images_in), boxes_in.float())
        return {"activities": scores}


class _ContextHead(nn.Module):
    """Parameter container of one EmbfeatureContextEncodingTransformer head, layer 1 (TCE_STBiP_module.py:224-250)."""

    def __init__(self, dim, nfb, dropout=0.1):
        super().__init__()
        self.downsample2 = nn.Conv2d(512, dim, kernel_size=1, stride=1)
        self.emb_roi = nn.Linear(nfb, dim, bias=True)
        self.dropout = nn.Dropout(dropout)
        self.layernorm1 = nn.LayerNorm(dim)
        self.FFN = nn.Sequential(nn.Linear(dim, dim, bias=True), nn.ReLU(inplace=True), nn.Dropout(dropout),
                                 nn.Linear(dim, dim, bias=True))
        self.layernorm2 = nn.LayerNorm(dim)


class _ContextEncoding(nn.Module):
    """MultiHeadLayerEmbfeatureContextEncoding(heads, 1 layer, ...) (TCE_STBiP_module.py:289-310): state_dict keys
    `CET.<h>.{downsample2, emb_roi, layernorm1, FFN.0, FFN.3, layernorm2}.{weight, bias}`.  Never called: the plan
    (din_b200.engine.TCEWeights) runs the four heads as one tcgen05 GEMM + the context attention kernel."""

    def __init__(self, heads, dim, nfb):
        super().__init__()
        self.CET = nn.ModuleList([_ContextHead(dim, nfb) for _ in range(heads)])


class Dynamic_TCE_volleyball(_DinModel):
    """reference infer_model.py:237-468: Dynamic_volleyball with the temporal-context-encoding module prepended to the
    dynamic inference; trains too (scripts/train_volleyball_stage2_dynamic_tce.py): din_context_attention_bwd_f32."""
    _dataset = "volleyball"
    _tce = True

    def __init__(self, cfg):
        super().__init__()
        self._common_init(cfg, (10, 12), lambda T, N, C: [T, N, C])

    def forward(self, batch_data):
        images_in, boxes_in = batch_data
        dev = self._check_mode(images_in)
        with torch.cuda.device(dev):
            if not boxes_in.is_cuda:
                boxes_in, = self._to_device(dev, boxes_in)
            scores = self._run(_as_frames(images_in), boxes_in.float())
        return {"activities": scores}


class Dynamic_collective(_DinModel):
    """reference infer_model.py:1135-1319 (variable actor count per clip, one launch instead of a loop)."""
    _dataset = "collective"

    def __init__(self, cfg):
        super().__init__()
        self._common_init(cfg, (cfg.num_frames, cfg.num_boxes), lambda T, N, C: [T, C])

    def forward(self, batch_data):
        images_in, boxes_in, bboxes_num_in = batch_data
        dev = self._check_mode(images_in)
        with torch.cuda.device(dev):
            if not (boxes_in.is_cuda and bboxes_num_in.is_cuda):
                boxes_in, bboxes_num_in = self._to_device(dev, boxes_in, bboxes_num_in)
            scores = self._run(_as_frames(images_in), boxes_in.float(), bboxes_num_in)
        return {"activities": scores}


def _out_of_scope(name):
    class _Stub(nn.Module):
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name} is outside the DIN stage-2 hot-path scope (SURVEY.md §2 #13/#14)")
    _Stub.__name__ = name
    return _Stub


# names the reference trainer's registry resolves at import time (train_net_dynamic.py:66-73)
PCTDM_volleyball = _out_of_scope("PCTDM_volleyball")
HiGCIN_volleyball = _out_of_scope("HiGCIN_volleyball")
AT_volleyball = _out_of_scope("AT_volleyball")
ARG_volleyball = _out_of_scope("ARG_volleyball")
SACRF_BiUTE_volleyball = _out_of_scope("SACRF_BiUTE_volleyball")
