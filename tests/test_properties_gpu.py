"""Size-independent properties of the CUDA path at BASELINE sizes (where the CPU oracle would take minutes):
determinism, clip independence / permutation equivariance (the forward pass shards over clips), linearity of
the tcgen05 convolution, and Collective's insensitivity to padded actor slots.  All through the public model
API / the C ABI; all comparisons bit-exact."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(cuda, pc, seed=0):
    import din_oracle as O
    import infer_model as IM
    from config import Config
    bb = O.build_backbone(pc.backbone)
    sd = O.make_state_dict(pc, seed=seed, backbone=bb)
    cfg = Config(pc.dataset)
    cfg.log_path = None
    for k in ("backbone", "image_size", "out_size", "emb_features", "num_frames", "num_boxes", "crop_size",
              "num_features_boxes", "num_activities", "lite_dim", "ST_kernel_size", "scale_factor", "beta_factor",
              "hierarchical_inference", "num_DIM"):
        setattr(cfg, k, getattr(pc, k))
    cfg.sampling_ratio = list(pc.sampling_ratio)
    m = (IM.Dynamic_collective if pc.dataset == "collective" else IM.Dynamic_volleyball)(cfg)
    m.load_state_dict(sd)
    return m.to(cuda).eval()


def test_headline_config_clip_independence_720p(cuda):
    """VGG-16 lite, T=10, N=12, 720x1280 (the bench workload): running 3 clips together, alone, or permuted
    gives bit-identical logits per clip, and a repeated call is bit-identical (no atomics, fixed K order)."""
    import din_oracle as O
    pc = O.PathConfig()                       # defaults = scripts/train_volleyball_stage2_dynamic.py
    model = _model(cuda, pc)
    images, boxes = O.make_inputs(pc, 3, seed=5)
    images, boxes = images.to(cuda), boxes.to(cuda)
    with torch.no_grad():
        a = model((images, boxes))["activities"]
        b = model((images, boxes))["activities"]
        assert torch.equal(a, b)
        assert torch.isfinite(a).all() and a.shape == (3, 8)
        for i in range(3):
            solo = model((images[i:i + 1], boxes[i:i + 1]))["activities"]
            assert torch.equal(solo[0], a[i]), i
        perm = torch.tensor([2, 0, 1], device=cuda)
        p = model((images[perm].contiguous(), boxes[perm].contiguous()))["activities"]
        assert torch.equal(p, a[perm])
    # logits actually depend on the input (guards against a constant output passing the checks above)
    assert (a[0] - a[1]).abs().max().item() > 1e-3


def test_conv_linearity_full_size(cuda):
    """conv(2x) == 2 conv(x) and conv(x, 2w) == 2 conv(x, w) exactly (powers of two commute with every fp16 /
    fp32 rounding), on VGG conv1_2's full-size shape through the HALO path."""
    from din_b200 import ops
    g = torch.Generator().manual_seed(21)
    x = torch.randn(2, 720, 1280, 64, generator=g).to(cuda).half()
    # weights kept away from the fp16 subnormal range, and fp32 outputs: only there is scaling by 2 exact
    r = torch.randn(64, 64, 3, 3, generator=g)
    w = (torch.sign(r) * (0.01 + r.abs() * 0.04)).to(cuda)
    y1 = ops.conv2d_nhwc(x, ops.pack_conv_weight(w), None, pad=(1, 1), out_f32=True)
    y2 = ops.conv2d_nhwc(x * 2, ops.pack_conv_weight(w), None, pad=(1, 1), out_f32=True)
    y3 = ops.conv2d_nhwc(x, ops.pack_conv_weight(w * 2), None, pad=(1, 1), out_f32=True)
    torch.cuda.synchronize()
    assert torch.equal(y2, y1 * 2)
    assert torch.equal(y3, y1 * 2)
    # zero input -> bias only, everywhere (padding is zero fill, not garbage)
    b = torch.randn(64, generator=g).to(cuda)
    y0 = ops.conv2d_nhwc(torch.zeros_like(x), ops.pack_conv_weight(w), b, pad=(1, 1))
    assert torch.equal(y0, b.half().view(1, 1, 1, 64).expand_as(y0))


def test_collective_ignores_padded_actors(cuda):
    """Collective (variable N): whatever sits in the boxes of padded actor slots (index >= bboxes_num) must not
    influence the logits — the reference slices them away (infer_model.py:1289-1290)."""
    import din_oracle as O
    pc = O.PathConfig(dataset="collective", backbone="res18", image_size=(480, 720), out_size=(15, 23), num_frames=10,
                      num_boxes=13, lite_dim=None, ST_kernel_size=(3, 3), num_activities=4)
    model = _model(cuda, pc)
    images, boxes, nb = O.make_inputs(pc, 4, seed=9)
    images, boxes, nb = images.to(cuda), boxes.to(cuda), nb.to(cuda)
    with torch.no_grad():
        a = model((images, boxes, nb))["activities"]
        junk = boxes.clone()
        for b in range(4):
            n = int(nb[b, 0])
            junk[b, :, n:, :] = torch.tensor([3.0, 2.0, 9.0, 11.0], device=cuda)
        c = model((images, junk, nb))["activities"]
    assert torch.equal(a, c)
    assert torch.isfinite(a).all() and a.shape == (4, 4)
