"""Edge cases of the path's input contract, CUDA path vs the CPU oracle: degenerate clip shapes (one frame, one
actor, one clip), image sizes that are not multiples of the backbone stride or of the kernels' tiles, boxes
entirely outside the map, Collective clips with a single real actor, and argument errors that must be reported
(not crash) by the C ABI."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(cuda, **kw):
    from test_e2e_gpu import _pc, _run_case
    backbone = kw.pop("backbone", "vgg16")
    hw = kw.pop("hw")
    B = kw.pop("B")
    tol = kw.pop("tol", 1e-3)
    return _run_case(cuda, _pc(backbone, hw, **kw), B, tol=tol)


def test_single_frame_single_actor_single_clip(cuda):
    # T = 1, N = 1: the 3x3 interaction field lies entirely in zero padding except its centre.  One actor in one
    # frame on a 2x3 map has nothing to average fp16 activation rounding over; this seed is the tail of the per-clip
    # error distribution (1.0e-3 .. 1.4e-3 under every precision setting, 7 of 8 seeds are under 1e-3: see
    # tests/test_fullsize_gpu.py::test_edge_degenerate_shapes_over_seeds and profiles/edge_precision_study_r2.md), so
    # this SHAPE check allows 2e-3; every BASELINE-shaped configuration is held to north_star's 1e-3.
    _case(cuda, hw=(64, 96), B=1, num_frames=1, num_boxes=1, tol=2e-3)


def test_one_frame_many_actors_and_many_frames_one_actor(cuda):
    _case(cuda, hw=(64, 96), B=2, num_frames=1, num_boxes=12)
    _case(cuda, hw=(64, 96), B=2, num_frames=10, num_boxes=1)


def test_image_size_not_multiple_of_stride_or_tile(cuda):
    # 75 x 109: VGG's five floor-mode pools drop odd rows/cols at several levels; output 2 x 3
    _case(cuda, hw=(75, 109), B=1, num_frames=2, num_boxes=3)
    # ResNet-18: odd extents through the 7x7 s2 stem, the 3x3 s2 pool and three stride-2 stages
    _case(cuda, backbone="res18", hw=(77, 115), B=1, num_frames=2, num_boxes=3)


def test_wide_image_many_strips(cuda):
    # width > 128*k: several stem strips and conv tiles per row, ragged last strip (1290 = 10*128 + 10)
    _case(cuda, hw=(64, 1290), B=1, num_frames=1, num_boxes=2)


def test_boxes_outside_the_map_and_bad_frame_index(cuda):
    """Boxes far outside the feature map crop to all zeros (extrapolation value); a box index outside
    [0, n_img) yields zeros instead of an out-of-bounds read."""
    import din_oracle as O
    from din_b200 import ops
    g = torch.Generator().manual_seed(3)
    fm = torch.randn(2, 16, 10, 14, generator=g).half().float()
    boxes = torch.tensor([[100.0, 100.0, 104.0, 107.0], [-50.0, -60.0, -45.0, -52.0], [2.0, 3.0, 6.0, 8.0],
                          [2.0, 3.0, 6.0, 8.0]])
    idx = torch.tensor([0, 1, 7, -1], dtype=torch.int32)
    out = ops.roi_align_nhwc(fm.permute(0, 2, 3, 1).contiguous().half().to(cuda), boxes.to(cuda), idx.to(cuda), 5, 5)
    assert (out == 0).all()
    ref = O.roi_align_longcw(fm, boxes, idx, 5, 5)
    assert (ref == 0).all()


def test_collective_single_real_actor(cuda):
    import din_oracle as O
    from test_e2e_gpu import _pc
    import infer_model as IM
    from config import Config
    pc = _pc("res18", (96, 144), dataset="collective", num_frames=3, num_boxes=13, lite_dim=None,
             ST_kernel_size=(3, 3), num_activities=4)
    bb = O.build_backbone(pc.backbone)
    sd = O.make_state_dict(pc, seed=4, backbone=bb)
    O.load_backbone(bb, sd)
    images, boxes, nb = O.make_inputs(pc, 3, seed=4)
    nb[:] = torch.tensor([1, 13, 2], dtype=torch.int32).view(3, 1)
    for b, n in enumerate((1, 13, 2)):
        boxes[b, :, n:, :] = 0
    ref = O.collective_forward(bb, sd, pc, images, boxes, nb)
    cfg = Config("collective")
    cfg.log_path = None
    for k in ("backbone", "image_size", "out_size", "emb_features", "num_frames", "num_boxes", "crop_size",
              "num_features_boxes", "num_activities", "lite_dim", "ST_kernel_size", "scale_factor", "beta_factor",
              "hierarchical_inference", "num_DIM"):
        setattr(cfg, k, getattr(pc, k))
    cfg.sampling_ratio = list(pc.sampling_ratio)
    m = IM.Dynamic_collective(cfg)
    m.load_state_dict(sd)
    m = m.to(cuda).eval()
    with torch.no_grad():
        out = m((images.to(cuda), boxes.to(cuda), nb.to(cuda)))["activities"].cpu()
    assert (out - ref).abs().max().item() <= 1e-3 * ref.abs().max().item()


def test_abi_reports_errors_instead_of_crashing(cuda):
    from din_b200 import _lib, ops
    x = torch.zeros(1, 8, 8, 64, device=cuda, dtype=torch.float16)
    w = ops.pack_conv_weight(torch.zeros(64, 64, 3, 3, device=cuda))
    with pytest.raises(_lib.DinError, match="stride"):
        ops.conv2d_nhwc(x, w, None, stride=3, pad=(1, 1))
    with pytest.raises(_lib.DinError, match="contiguous CUDA"):
        ops.conv2d_nhwc(x.cpu(), w, None, pad=(1, 1))
    with pytest.raises(_lib.DinError, match="n <="):
        ops.dynamic_infer(torch.zeros(1, 2, 17, 64, device=cuda), torch.zeros(9, 27, 64, device=cuda),
                          torch.zeros(27, device=cuda), (3, 3), 1)
    # the models refuse CPU tensors and training a backbone without backward kernels (no silent fallback); training
    # the head, and the VGG-16 backbone, is supported (tests/test_backward_gpu.py)
    import infer_model as IM
    from config import Config
    cfg = Config("volleyball")
    cfg.log_path = None
    cfg.backbone, cfg.out_size, cfg.emb_features, cfg.image_size = "vgg16", (2, 3), 512, (64, 96)
    cfg.ST_kernel_size, cfg.sampling_ratio, cfg.beta_factor, cfg.lite_dim = [(3, 3)], [1], False, 128
    m = IM.Dynamic_volleyball(cfg)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.eval()((torch.zeros(1, 3, 3, 64, 96), torch.zeros(1, 3, 12, 4)))
    cfg.train_backbone, cfg.backbone, cfg.out_size, cfg.emb_features = True, "inv3", (2, 3), 1056
    m = IM.Dynamic_volleyball(cfg).to(cuda).train()          # BatchNorm left on batch statistics (ResNet-18 only)
    with pytest.raises(NotImplementedError, match="BatchNorm"):
        m((torch.zeros(1, 3, 3, 64, 96, device=cuda), torch.zeros(1, 3, 12, 4, device=cuda)))
