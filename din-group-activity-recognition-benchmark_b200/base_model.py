"""Drop-in for the stage-1 base models of `base_model` (reference base_model.py:6-284).

`Basenet_volleyball(cfg)` / `Basenet_collective(cfg)` keep the reference's constructor contract, module
tree, state_dict key names, `savemodel` / `loadmodel` checkpoint format (the file `Dynamic_*.loadmodel`
reads in stage 2: keys `backbone_state_dict`, `fc_emb_state_dict`, base_model.py:46-55,181-190) and
`forward(batch) -> (actions_scores, activities_scores)`.  forward() runs the same sm_100a kernels as the
stage-2 path (backbone plan, RoIAlign, tcgen05 embedding GEMM with the ReLU fused) plus the two small heads;
torch modules below are parameter containers and are never called.

Scope: evaluation / inference forward; `Basenet_volleyball` with the VGG-16 backbone -- the
model scripts/train_volleyball_stage1.py trains -- also runs its training step on the CUDA path (one autograd node:
both heads, dropout, the embedding and the whole backbone).  Other stage-1 training
configurations raise.
"""
import torch
import torch.nn as nn

from backbone.backbone import MyInception_v3, MyRes18, MyVGG16
from din_b200 import plan_cache as _pc
from din_b200 import train as _train
from din_b200.engine import BasenetEngine
from roi_align.roi_align import RoIAlign


class _BasenetTrainFn(torch.autograd.Function):
    """Stage-1 forward + backward as one autograd node (din_b200.train.basenet_forward_train / basenet_backward)."""

    @staticmethod
    def forward(ctx, model, images, boxes, names, *params):
        train_bb = any(n.startswith("backbone.") for n in names)
        (actions, activities), tape = _train.basenet_forward_train(model.engine(), images, boxes,
                                                                    training=model.training, train_backbone=train_bb)
        ctx.eng, ctx.tape, ctx.names = model.engine(), tape, names
        ctx.sink = getattr(_pc.owner_of(model), "grad_sink", None)
        ctx.shapes = [tuple(p.shape) for p in params]
        return actions, activities

    @staticmethod
    def backward(ctx, dactions, dactivities):
        sink = ctx.sink if (ctx.sink is not None and ctx.sink.active()) else None
        grads = _train.basenet_backward(ctx.eng, ctx.tape, dactions, dactivities, sink=sink)
        ctx.tape = None
        if sink is not None:
            sink.finish()
            return (None,) * (4 + len(ctx.names))
        out = [grads[n].reshape(shp) if (need and n in grads) else None
               for n, shp, need in zip(ctx.names, ctx.shapes, ctx.needs_input_grad[4:])]
        return (None, None, None, None) + tuple(out)


class _Basenet(nn.Module):
    _dataset = None
    _emb_name = None

    def _heads(self, cfg):
        D, K, NFB = cfg.emb_features, cfg.crop_size[0], cfg.num_features_boxes
        self.roi_align = RoIAlign(*cfg.crop_size)
        setattr(self, self._emb_name, nn.Linear(K * K * D, NFB))
        self.fc_actions = nn.Linear(NFB, cfg.num_actions)
        self.fc_activities = nn.Linear(NFB, cfg.num_activities)
        self._owner = [self]               # nn.DataParallel replicas find the original model here (din_b200/plan_cache.py)
        self._plans = _pc.PlanTable()

    def savemodel(self, filepath):
        state = {
            "backbone_state_dict": self.backbone.state_dict(),
            "fc_emb_state_dict": getattr(self, self._emb_name).state_dict(),
            "fc_actions_state_dict": self.fc_actions.state_dict(),
            "fc_activities_state_dict": self.fc_activities.state_dict(),
        }
        torch.save(state, filepath)
        print("model saved to:", filepath)

    def engine(self):
        own, dev, key = _pc.owner_of(self), _pc.device_of(self), _pc.version_key(self)
        slot = own._plans.get(dev)
        if slot is None or slot["key"] != key:
            _pc.require_cuda(dev)
            with torch.cuda.device(dev):
                eng = BasenetEngine(self.cfg, _pc.named_tensors(self), dev, dataset=self._dataset,
                                    emb_name=self._emb_name)
            slot = own._plans.put(dev, key=key, engine=eng)
        return slot["engine"]

    def _check_mode(self, images):
        if not images.is_cuda:
            raise RuntimeError("the DIN hot path runs on sm_100a only: pass CUDA tensors (there is no CPU fallback)")
        if self.training and (self._dataset != "volleyball" or self.cfg.backbone not in ("vgg16", "res18", "inv3")):
            raise NotImplementedError(
                "stage-1 training on the sm_100a path is implemented for Basenet_volleyball with the VGG-16, ResNet-18 or "
                "Inception-v3 backbone (scripts/train_volleyball_stage1.py); use model.eval() for Basenet_collective "
                "(SURVEY.md §8f rank 1)")
        if self.training and any(isinstance(m, nn.modules.batchnorm._BatchNorm) and m.training
                                 for m in self.backbone.modules()):
            raise NotImplementedError(
                "stage-1 training with BatchNorm on batch statistics is not implemented on the sm_100a path (stage 2 / "
                "infer_model.py has it for ResNet-18): freeze BN as cfg.set_bn_eval does (train_net.py model.apply(set_bn_eval))")


class Basenet_volleyball(_Basenet):
    """reference base_model.py:6-142."""
    _dataset, _emb_name = "volleyball", "fc_emb"

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        if cfg.backbone == "inv3":
            self.backbone = MyInception_v3(transform_input=False, pretrained=True)
        elif cfg.backbone == "vgg16":
            self.backbone = MyVGG16(pretrained=True)
        elif cfg.backbone == "res18":
            self.backbone = MyRes18(pretrained=True)
        else:
            raise NotImplementedError(f"backbone {cfg.backbone!r} is outside the DIN hot-path scope")
        self._heads(cfg)
        self.dropout_emb = nn.Dropout(p=cfg.train_dropout_prob)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.kaiming_normal_(m.weight)
                nn.init.zeros_(m.bias)

    def loadmodel(self, filepath):
        state = torch.load(filepath)
        self.backbone.load_state_dict(state["backbone_state_dict"])
        self.fc_emb.load_state_dict(state["fc_emb_state_dict"])
        self.fc_actions.load_state_dict(state["fc_actions_state_dict"])
        self.fc_activities.load_state_dict(state["fc_activities_state_dict"])
        print("Load model states from: ", filepath)

    def forward(self, batch_data):
        images_in, boxes_in = batch_data
        self._check_mode(images_in)
        frames = images_in if images_in.dtype == torch.uint8 else images_in.float()
        with torch.cuda.device(images_in.device):
            if not self.training:
                return self.engine().forward_volleyball(frames, boxes_in.float())
            if not torch.is_grad_enabled():
                return _train.basenet_forward_train(self.engine(), frames, boxes_in.float(), training=True,
                                                    train_backbone=False)[0]
            named = _pc.trainable(self)
            return _BasenetTrainFn.apply(self, frames, boxes_in.float(), tuple(n for n, _ in named),
                                         *[p for _, p in named])


class Basenet_collective(_Basenet):
    """reference base_model.py:145-284 (Inception-v3 hard-coded :159; actor count varies per frame)."""
    _dataset, _emb_name = "collective", "fc_emb_1"

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.backbone = MyInception_v3(transform_input=False, pretrained=True)
        if not cfg.train_backbone:
            for p in self.backbone.parameters():
                p.requires_grad = False
        self._heads(cfg)
        self.dropout_emb_1 = nn.Dropout(p=cfg.train_dropout_prob)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.kaiming_normal_(m.weight)

    def loadmodel(self, filepath):
        state = torch.load(filepath)
        self.backbone.load_state_dict(state["backbone_state_dict"])
        self.fc_emb_1.load_state_dict(state["fc_emb_state_dict"])
        print("Load model states from: ", filepath)

    def forward(self, batch_data):
        images_in, boxes_in, bboxes_num_in = batch_data
        self._check_mode(images_in)
        frames = images_in if images_in.dtype == torch.uint8 else images_in.float()
        with torch.cuda.device(images_in.device):
            return self.engine().forward_collective(frames, boxes_in.float(), bboxes_num_in)
