"""make_golden.py — generates tests/golden/*.pt from the REFERENCE's own classes.  TEST INFRASTRUCTURE.

Run in the authoring container (needs /root/reference):   python oracle/make_golden.py
Each fixture stores the PathConfig, batch size and seed (weights and inputs are regenerated from the seed
by din_oracle.make_state_dict / make_inputs, with float64 checksums recorded to detect RNG drift) and the
logits the reference model (infer_model.Dynamic_volleyball / Dynamic_collective, patched per
ref_harness.py where it crashes as shipped) produced.  Module-level fixtures store full tensors.
"""
import dataclasses
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import din_oracle as O  # noqa: E402
import ref_harness as R  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def checksum(tensors):
    return float(sum(t.double().abs().sum() for t in tensors if t.is_floating_point()))


def model_cases():
    def pc(backbone, hw, **kw):
        return O.PathConfig(backbone=backbone, image_size=hw, out_size=O.backbone_out_size(backbone, *hw), **kw)
    return {
        "vgg16_lite": (pc("vgg16", (96, 160), num_frames=3, num_boxes=4), 2),
        "res18_lite": (pc("res18", (96, 160), num_frames=3, num_boxes=4), 2),
        "vgg16_full_r13_beta": (pc("vgg16", (96, 160), num_frames=4, num_boxes=5, lite_dim=None,
                                   sampling_ratio=(1, 3), beta_factor=True), 2),
        "vgg16_parallel_fields": (pc("vgg16", (96, 160), num_frames=4, num_boxes=5,
                                     ST_kernel_size=[(1, 3), (3, 1)], num_DIM=2), 2),
        "vgg16_hierarchical": (pc("vgg16", (64, 96), num_frames=10, num_boxes=12, lite_dim=None,
                                  ST_kernel_size=[(1, 3), (3, 1)], hierarchical_inference=True), 1),
        "inv3_full": (pc("inv3", (139, 203), emb_features=1056, num_frames=2, num_boxes=4, lite_dim=None), 2),
        "collective_res18": (pc("res18", (96, 144), dataset="collective", num_frames=3, num_boxes=13,
                                lite_dim=None, ST_kernel_size=(3, 3), num_activities=4), 3),
        # Dynamic_TCE_volleyball (infer_model.py:237-468): context encoding prepended to DIN; N = 12 is asserted (:261)
        "tce_vgg16": (pc("vgg16", (96, 160), num_frames=3, num_boxes=12, lite_dim=None, tce=True), 2),
    }


def basenet_cases():
    def pc(backbone, hw, **kw):
        return O.PathConfig(backbone=backbone, image_size=hw, out_size=O.backbone_out_size(backbone, *hw), **kw)
    return {
        "volleyball_vgg16": (pc("vgg16", (96, 160), num_frames=3, num_boxes=4), 2),
        "volleyball_res18_T1": (pc("res18", (96, 160), num_frames=1, num_boxes=4), 2),
        "collective_inv3": (pc("inv3", (139, 203), dataset="collective", emb_features=1056, num_frames=2,
                               num_boxes=13, num_activities=5, num_actions=6), 2),
    }


def main_basenet():
    """Stage-1 fixtures from the reference's base_model.Basenet_* (SURVEY.md §8f rank 3)."""
    for name, (pc, B) in basenet_cases().items():
        bb = O.build_backbone(pc.backbone)
        sd = O.make_basenet_state_dict(pc, seed=0, backbone=bb)
        batch = O.make_basenet_inputs(pc, B, seed=0)
        actions, activities = R.ref_basenet_forward(pc, sd, *batch)
        torch.save({"config": dataclasses.asdict(pc), "B": B, "seed": 0, "actions_ref": actions,
                    "activities_ref": activities, "weights_checksum": checksum(sd.values()),
                    "inputs_checksum": checksum(batch)}, os.path.join(OUT, f"basenet_{name}.pt"))
        print("basenet", name, tuple(actions.shape), tuple(activities.shape))


def grad_cases():
    c = model_cases()
    return {k: c[k] for k in ("vgg16_lite", "vgg16_full_r13_beta", "vgg16_parallel_fields", "collective_res18")}


def main_grads():
    """Training-step fixtures: the REFERENCE model's own autograd gradients (ref_harness.ref_head_grads)."""
    for name, (pc, B) in grad_cases().items():
        bb = O.build_backbone(pc.backbone)
        sd = O.make_state_dict(pc, seed=0, backbone=bb)
        batch = O.make_inputs(pc, B, seed=0)
        labels = torch.arange(B) % pc.num_activities
        logits, loss, grads = R.ref_head_grads(pc, sd, labels, *batch)
        torch.save({"config": dataclasses.asdict(pc), "B": B, "seed": 0, "labels": labels, "logits_ref": logits,
                    "loss_ref": loss, "grads_ref": {k: O.grad_digest(v) for k, v in grads.items()},
                    "weights_checksum": checksum(sd.values()), "inputs_checksum": checksum(batch)},
                   os.path.join(OUT, f"grads_{name}.pt"))
        print("grads", name, float(loss), len(grads))
    main_fullgrads()


def main_fullgrads():
    """The whole model trained, backbone included (scripts/train_volleyball_stage2_dynamic.py:12); ResNet-18 with its
    BatchNorm layers in eval mode (config.py set_bn_eval / train_net.py:25-28)."""
    names = ("vgg16_lite", "res18_lite", "collective_res18", "inv3_full")   # the third: scripts/train_collective_stage2_dynamic.py
    if "--missing-only" in sys.argv:
        names = tuple(n for n in names if not os.path.exists(os.path.join(OUT, f"fullgrads_{n}.pt")))
    for name in names:
        pc, B = model_cases()[name]
        bb = O.build_backbone(pc.backbone)
        sd = O.make_state_dict(pc, seed=0, backbone=bb)
        batch = O.make_inputs(pc, B, seed=0)
        labels = torch.arange(B) % pc.num_activities
        logits, loss, grads = R.ref_head_grads(pc, sd, labels, *batch, train_backbone=True)
        torch.save({"config": dataclasses.asdict(pc), "B": B, "seed": 0, "labels": labels, "logits_ref": logits,
                    "loss_ref": loss, "grads_ref": {k: O.grad_digest(v) for k, v in grads.items()},
                    "weights_checksum": checksum(sd.values()), "inputs_checksum": checksum(batch),
                    "train_backbone": True}, os.path.join(OUT, f"fullgrads_{name}.pt"))
        print("fullgrads", name, float(loss), len(grads))
    # BatchNorm on batch statistics (cfg.set_bn_eval = False, the default: scripts/train_collective_stage2_dynamic.py)
    for name in ("res18_lite", "collective_res18"):
        if "--missing-only" in sys.argv and os.path.exists(os.path.join(OUT, f"bntrain_{name}.pt")):
            continue
        pc, B = model_cases()[name]
        bb = O.build_backbone(pc.backbone)
        sd = O.make_state_dict(pc, seed=0, backbone=bb)
        batch = O.make_inputs(pc, B, seed=0)
        labels = torch.arange(B) % pc.num_activities
        logits, loss, grads, bufs = R.ref_head_grads(pc, sd, labels, *batch, train_backbone=True, bn_train=True,
                                                     return_buffers=True)
        torch.save({"config": dataclasses.asdict(pc), "B": B, "seed": 0, "labels": labels, "logits_ref": logits,
                    "loss_ref": loss, "grads_ref": {k: O.grad_digest(v) for k, v in grads.items()},
                    "buffers_ref": {k: O.grad_digest(v.float()) for k, v in bufs.items()},
                    "weights_checksum": checksum(sd.values()), "inputs_checksum": checksum(batch),
                    "train_backbone": True, "bn_train": True}, os.path.join(OUT, f"bntrain_{name}.pt"))
        print("bntrain", name, float(loss), len(grads), len(bufs))


def main_basenet_grads():
    """Stage-1 training-step fixture from the reference's own Basenet_volleyball (VGG-16, T = 1 as
    scripts/train_volleyball_stage1.py, class-weighted action loss)."""
    pc = O.PathConfig(backbone="vgg16", image_size=(96, 160), out_size=O.backbone_out_size("vgg16", 96, 160),
                      num_frames=1, num_boxes=4)
    B = 3
    bb = O.build_backbone(pc.backbone)
    sd = O.make_basenet_state_dict(pc, seed=0, backbone=bb)
    batch = O.make_basenet_inputs(pc, B, seed=0)
    g = torch.Generator().manual_seed(9)
    actions_labels = torch.randint(0, pc.num_actions, (B * pc.num_boxes,), generator=g)
    activities_labels = torch.randint(0, pc.num_activities, (B,), generator=g)
    weights = torch.tensor([1., 1., 2., 3., 1., 2., 2., 0.2, 1.])       # scripts/train_volleyball_stage1.py:33
    loss, grads = R.ref_basenet_grads(pc, sd, actions_labels, activities_labels, weights, *batch)
    torch.save({"config": dataclasses.asdict(pc), "B": B, "seed": 0, "actions_labels": actions_labels,
                "activities_labels": activities_labels, "actions_weights": weights, "loss_ref": loss,
                "grads_ref": {k: O.grad_digest(v) for k, v in grads.items()},
                "weights_checksum": checksum(sd.values())}, os.path.join(OUT, "stage1grads_vgg16_T1.pt"))
    print("basenet grads", float(loss), len(grads))


def main():
    os.makedirs(OUT, exist_ok=True)
    if "--fullgrads-only" in sys.argv:
        main_fullgrads()
        sys.exit(0)
    if "--basenet-grads-only" in sys.argv:
        return main_basenet_grads()
    if "--grads-only" in sys.argv:
        return main_grads()
    if "--basenet-only" in sys.argv:
        return main_basenet()
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None      # one whole-model fixture
    if only is None:
        main_basenet()
        main_grads()
        main_basenet_grads()
    for name, (pc, B) in model_cases().items():
        if only is not None and name != only:
            continue
        bb = O.build_backbone(pc.backbone)
        sd = O.make_state_dict(pc, seed=0, backbone=bb)
        batch = O.make_inputs(pc, B, seed=0)
        logits = R.ref_forward(pc, sd, *batch)
        torch.save({"config": dataclasses.asdict(pc), "B": B, "seed": 0, "logits_ref": logits,
                    "weights_checksum": checksum(sd.values()), "inputs_checksum": checksum(batch)},
                   os.path.join(OUT, f"model_{name}.pt"))
        print(name, logits[0, :4].tolist())
    if only is not None:
        return
    # module-level: the reference Dynamic_Person_Inference with randomised p_conv / scale_conv
    g = torch.Generator().manual_seed(42)
    for name, (kernel, ratios, T, N, beta) in {
        "dpi_k33_r1": ((3, 3), [1], 4, 5, False),
        "dpi_k33_r13_beta": ((3, 3), [1, 3], 10, 12, True),
        "dpi_k13_r1": ((1, 3), [1], 4, 5, False),
        "dpi_k31_r2": ((3, 1), [2], 4, 5, False),
    }.items():
        C = 32
        m = R.ref_dpi_module(C, kernel, ratios, scale_factor=True, beta_factor=beta)
        with torch.no_grad():
            for n, p in m.named_parameters():
                if "p_conv" in n or "scale_conv" in n:
                    p.copy_(torch.randn(p.shape, generator=g) * (0.02 if n.endswith("weight") else 1.5))
                elif n == "beta":
                    p.copy_(torch.rand(p.shape, generator=g) + 0.5)
        x = torch.randn(2, T, N, C, generator=g)
        with torch.no_grad():
            y, mad = m(x)
        torch.save({"kernel": kernel, "ratios": ratios, "beta": beta, "x": x, "y": y,
                    "state_dict": {k: v.clone() for k, v in m.state_dict().items()}},
                   os.path.join(OUT, f"module_{name}.pt"))
        print(name, tuple(y.shape))


if __name__ == "__main__":
    main()
