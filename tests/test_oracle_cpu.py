"""CPU tests (run everywhere): the oracle restatement against (a) the committed golden fixtures produced by
the reference's own classes and (b) the live reference when /root/reference is present."""
import glob
import os

import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _checksum(tensors):
    return float(sum(t.double().abs().sum() for t in tensors if t.is_floating_point()))


def _pc_from(d):
    import din_oracle as O
    d = dict(d)
    ks = d["ST_kernel_size"]
    d["ST_kernel_size"] = tuple(ks) if isinstance(ks[0], int) else [tuple(k) for k in ks]
    for k in ("image_size", "out_size", "crop_size", "sampling_ratio"):
        d[k] = tuple(d[k])
    return O.PathConfig(**d)


MODEL_FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "model_*.pt")))
MODULE_FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "module_*.pt")))
BASENET_FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "basenet_*.pt")))


def test_fixtures_present():
    assert len(MODEL_FIXTURES) == 8 and len(MODULE_FIXTURES) == 4 and len(BASENET_FIXTURES) == 3
    assert len(glob.glob(os.path.join(GOLDEN, "grads_*.pt"))) == 4


@pytest.mark.parametrize("path", BASENET_FIXTURES, ids=[os.path.basename(p) for p in BASENET_FIXTURES])
def test_oracle_reproduces_reference_basenet(path):
    """Stage-1 restatement == what base_model.Basenet_* produced (fixtures from oracle/make_golden.py)."""
    import din_oracle as O
    fx = torch.load(path)
    pc = _pc_from(fx["config"])
    bb = O.build_backbone(pc.backbone)
    sd = O.make_basenet_state_dict(pc, seed=fx["seed"], backbone=bb)
    batch = O.make_basenet_inputs(pc, fx["B"], seed=fx["seed"])
    assert abs(_checksum(sd.values()) - fx["weights_checksum"]) <= 1e-9 * fx["weights_checksum"]
    assert abs(_checksum(batch) - fx["inputs_checksum"]) <= 1e-9 * fx["inputs_checksum"]
    O.load_backbone(bb, sd)
    fwd = O.basenet_collective_forward if pc.dataset == "collective" else O.basenet_volleyball_forward
    actions, activities = fwd(bb, sd, pc, *batch)
    for out, ref in ((actions, fx["actions_ref"]), (activities, fx["activities_ref"])):
        assert out.shape == ref.shape
        assert (out - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()


@pytest.mark.parametrize("path", MODEL_FIXTURES, ids=[os.path.basename(p) for p in MODEL_FIXTURES])
def test_oracle_reproduces_reference_logits(path):
    """Oracle forward == logits the reference model produced (fp32 CPU both: 1e-5 * max|ref|)."""
    import din_oracle as O
    fx = torch.load(path)
    pc = _pc_from(fx["config"])
    bb = O.build_backbone(pc.backbone)
    sd = O.make_state_dict(pc, seed=fx["seed"], backbone=bb)
    batch = O.make_inputs(pc, fx["B"], seed=fx["seed"])
    # guard: the seeded weights / inputs are the ones the fixture was generated with
    assert abs(_checksum(sd.values()) - fx["weights_checksum"]) <= 1e-9 * fx["weights_checksum"]
    assert abs(_checksum(batch) - fx["inputs_checksum"]) <= 1e-9 * fx["inputs_checksum"]
    O.load_backbone(bb, sd)
    fwd = O.collective_forward if pc.dataset == "collective" else O.volleyball_forward
    out = fwd(bb, sd, pc, *batch)
    ref = fx["logits_ref"]
    assert (out - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()


@pytest.mark.parametrize("path", MODULE_FIXTURES, ids=[os.path.basename(p) for p in MODULE_FIXTURES])
def test_oracle_dpi_reproduces_reference_module(path):
    import din_oracle as O
    fx = torch.load(path)
    y = O.dynamic_person_inference(fx["x"], fx["state_dict"], "", tuple(fx["kernel"]), fx["ratios"], True,
                                   fx["beta"])
    assert (y - fx["y"]).abs().max().item() <= 1e-5 * fx["y"].abs().max().item()


def test_din_zero_init_is_window_mean():
    """Reference init (p_conv / scale_conv zero, dynamic_infer_module.py:66-67,80-81): offsets 0, relation
    1/k², lt weight 1 => DIN == zero-padded k-window mean."""
    import din_oracle as O
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 4, 5, 16, generator=g)
    C = 16
    z = lambda *s: torch.zeros(*s)
    out, _ = O.din_ratio(x, z(18, C, 3, 3), z(18), z(9, C, 3, 3), z(9), (3, 3), 1)
    ref = F.avg_pool2d(x.permute(0, 3, 1, 2), 3, 1, 1, count_include_pad=True).permute(0, 2, 3, 1)
    assert (out - ref).abs().max().item() < 1e-6


def test_din_border_double_count_quirk():
    """k_t = 1 => no padding along T: a sample pushed past the last frame collapses l and r onto it and is
    counted twice (dynamic_infer_module.py:220-233).  The oracle must reproduce the quirk."""
    import din_oracle as O
    C = 4
    x = torch.arange(1 * 3 * 2 * C, dtype=torch.float32).reshape(1, 3, 2, C) + 1
    p_w, s_w = torch.zeros(2, C, 1, 1), torch.zeros(1, C, 1, 1)
    p_b = torch.tensor([100.0, 0.0])          # T-axis offset far beyond the border, N-axis offset 0
    out, _ = O.din_ratio(x, p_w, p_b, s_w, torch.zeros(1), (1, 1), 1)
    # T axis: l = r = p = T-1 -> weight 1 each => every sample is the LAST frame, counted twice.
    # N axis (k_n = 1 => no padding either): actor 0 has l = p = 0, r = 1 (weight 0) -> x1;
    # actor 1 sits on the border: r clamps onto l -> counted twice again -> x2.  Total x2 and x4.
    last = x[:, 2:3].expand_as(out)
    assert torch.allclose(out[:, :, 0], 2 * last[:, :, 0])
    assert torch.allclose(out[:, :, 1], 4 * last[:, :, 1])


def test_roi_align_restated_semantics():
    """crop_and_resize facts the restatement must satisfy: bin centres sampled, exact at integer points,
    zero outside [0, H-1] x [0, W-1], all-zero for Collective's (0,0,0,0) padding boxes except the sample
    that lands inside the map."""
    import din_oracle as O
    H, W = 6, 8
    fm = (torch.arange(H).view(H, 1) * 10 + torch.arange(W).view(1, W)).float().view(1, 1, H, W)
    # box [1, 6] x [1, 6] with 5 bins of width 1: centres at 1.5 - 0.5 = 1, 2, ... => integer sample points
    out = O.roi_align_longcw(fm, torch.tensor([[1.0, 1.0, 6.0, 6.0]]), torch.tensor([0], dtype=torch.int32), 5, 5)
    assert torch.allclose(out[0, 0], fm[0, 0, 1:6, 1:6])
    out = O.roi_align_longcw(fm, torch.tensor([[-4.0, -4.0, 1.0, 1.0]]), torch.tensor([0], dtype=torch.int32), 5, 5)
    assert out[0, 0, :4].abs().sum() == 0 and out[0, 0, :, :4].abs().sum() == 0
    assert out[0, 0, 4, 4] == fm[0, 0, 0, 0]
    out = O.roi_align_longcw(fm, torch.tensor([[0.0, 0.0, 0.0, 0.0]]), torch.tensor([0], dtype=torch.int32), 5, 5)
    assert out.abs().sum() == 0            # sample point (-0.5, -0.5) is outside the map


def test_oracle_vs_live_reference():
    """Pin the oracle against the reference itself where it is available (authoring container only)."""
    import din_oracle as O
    import ref_harness as R
    if not R.available():
        pytest.skip("/root/reference not present on this machine (fixtures in tests/golden/ pin the oracle)")
    pc = O.PathConfig(backbone="res18", image_size=(64, 96), out_size=O.backbone_out_size("res18", 64, 96),
                      num_frames=2, num_boxes=3, sampling_ratio=(1, 2), beta_factor=True)
    bb = O.build_backbone(pc.backbone)
    sd = O.make_state_dict(pc, seed=3, backbone=bb)
    O.load_backbone(bb, sd)
    batch = O.make_inputs(pc, 2, seed=3)
    ref = R.ref_forward(pc, sd, *batch)
    out = O.volleyball_forward(bb, sd, pc, *batch)
    assert (out - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()


def test_config_defaults_match_reference():
    """The drop-in Config exposes exactly the reference Config's attribute names and default values."""
    import ref_harness as R
    if not R.available():
        pytest.skip("/root/reference not present on this machine")
    ref_cfg = R.ref_module("config").Config
    from config import Config
    for ds in ("volleyball", "collective"):
        a, b = vars(ref_cfg(ds)), vars(Config(ds))
        assert set(a) == set(b), set(a) ^ set(b)
        for k in a:
            assert a[k] == b[k], (ds, k, a[k], b[k])


GRAD_FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "grads_*.pt")))


def _digest_close(g, d, rtol=2e-4):
    """gradient tensor vs a grad_digest fixture: sampled values and norms, relative to max|grad|."""
    assert tuple(g.shape) == tuple(d["shape"]), (tuple(g.shape), d["shape"])
    flat = g.detach().double().flatten()
    scale = max(d["max_abs"], 1e-30)
    assert (flat[d["idx"]].float() - d["samples"]).abs().max().item() <= rtol * scale
    assert abs(float(flat.pow(2).sum().sqrt()) - d["l2"]) <= rtol * max(d["l2"], 1e-30)
    assert abs(float(flat.sum()) - d["sum"]) <= rtol * max(d["abs_sum"], 1e-30)


@pytest.mark.parametrize("path", GRAD_FIXTURES, ids=[os.path.basename(p) for p in GRAD_FIXTURES])
def test_oracle_autograd_reproduces_reference_gradients(path):
    """Training step: autograd over the restatement == the reference model's own
    gradients (frozen backbone, BN eval, dropout 0) for every parameter after the backbone."""
    import din_oracle as O
    fx = torch.load(path)
    pc = _pc_from(fx["config"])
    bb = O.build_backbone(pc.backbone)
    sd = O.make_state_dict(pc, seed=fx["seed"], backbone=bb)
    batch = O.make_inputs(pc, fx["B"], seed=fx["seed"])
    assert abs(_checksum(sd.values()) - fx["weights_checksum"]) <= 1e-9 * fx["weights_checksum"]
    O.load_backbone(bb, sd)
    bb.eval()
    logits, loss, grads = O.head_grads(bb, sd, pc, fx["labels"], *batch)
    assert (logits - fx["logits_ref"]).abs().max().item() <= 1e-5 * fx["logits_ref"].abs().max().item()
    assert abs(float(loss) - float(fx["loss_ref"])) <= 1e-5
    assert set(grads) == set(fx["grads_ref"]), set(grads) ^ set(fx["grads_ref"])
    for k, d in fx["grads_ref"].items():
        _digest_close(grads[k], d)


def test_oracle_autograd_reproduces_reference_backbone_gradients():
    """cfg.train_backbone = True (scripts/train_volleyball_stage2_dynamic.py:12): autograd over the restatement,
    backbone included, == the reference model's own gradients for all 43 parameter tensors."""
    import din_oracle as O
    fx = torch.load(os.path.join(GOLDEN, "fullgrads_vgg16_lite.pt"))
    pc = _pc_from(fx["config"])
    bb = O.build_backbone(pc.backbone)
    sd = O.make_state_dict(pc, seed=fx["seed"], backbone=bb)
    batch = O.make_inputs(pc, fx["B"], seed=fx["seed"])
    O.load_backbone(bb, sd)
    bb.eval()
    logits, loss, grads = O.head_grads(bb, sd, pc, fx["labels"], *batch, train_backbone=True)
    assert abs(float(loss) - float(fx["loss_ref"])) <= 1e-5
    assert set(grads) == set(fx["grads_ref"]), set(grads) ^ set(fx["grads_ref"])
    assert sum(k.startswith("backbone.") for k in grads) == 26
    for k, d in fx["grads_ref"].items():
        _digest_close(grads[k], d)


def test_oracle_autograd_reproduces_reference_stage1_gradients():
    """Stage-1 training step (train_net.py:163-189) on Basenet_volleyball / VGG-16: autograd over the restatement ==
    the reference model's own gradients (32 tensors: backbone, fc_emb, fc_actions, fc_activities)."""
    import din_oracle as O
    fx = torch.load(os.path.join(GOLDEN, "stage1grads_vgg16_T1.pt"))
    pc = _pc_from(fx["config"])
    bb = O.build_backbone(pc.backbone)
    sd = O.make_basenet_state_dict(pc, seed=fx["seed"], backbone=bb)
    assert abs(_checksum(sd.values()) - fx["weights_checksum"]) <= 1e-9 * fx["weights_checksum"]
    O.load_backbone(bb, sd)
    bb.eval()
    batch = O.make_basenet_inputs(pc, fx["B"], seed=fx["seed"])
    loss, grads = O.basenet_grads(bb, sd, pc, fx["actions_labels"], fx["activities_labels"], *batch,
                                  actions_weights=fx["actions_weights"])
    assert abs(float(loss) - float(fx["loss_ref"])) <= 1e-5
    assert set(grads) == set(fx["grads_ref"]), set(grads) ^ set(fx["grads_ref"])
    for k, d in fx["grads_ref"].items():
        _digest_close(grads[k], d)


@pytest.mark.parametrize("fixture,bn_train", [("fullgrads_res18_lite", False), ("fullgrads_collective_res18", False),
                                              ("bntrain_res18_lite", True), ("bntrain_collective_res18", True)])
def test_oracle_autograd_reproduces_reference_resnet_gradients(fixture, bn_train):
    """ResNet-18 trained (scripts/train_collective_stage2_dynamic.py:12-16), BatchNorm in eval mode (cfg.set_bn_eval) or
    on batch statistics (the config.py:80 default): autograd over the restatement == the reference model's own
    gradients, and -- batch statistics -- the same running statistics after the step."""
    import din_oracle as O
    fx = torch.load(os.path.join(GOLDEN, fixture + ".pt"))
    pc = _pc_from(fx["config"])
    bb = O.build_backbone(pc.backbone)
    sd = O.make_state_dict(pc, seed=fx["seed"], backbone=bb)
    batch = O.make_inputs(pc, fx["B"], seed=fx["seed"])
    O.load_backbone(bb, sd)
    bb.train(bn_train)
    logits, loss, grads = O.head_grads(bb, sd, pc, fx["labels"], *batch, train_backbone=True)
    assert abs(float(loss) - float(fx["loss_ref"])) <= 1e-5
    assert set(grads) == set(fx["grads_ref"]), set(grads) ^ set(fx["grads_ref"])
    for k, d in fx["grads_ref"].items():
        _digest_close(grads[k], d)
    if bn_train:
        bufs = {"backbone." + n: b for n, b in bb.named_buffers()}
        assert set(bufs) == set(fx["buffers_ref"])
        for k, d in fx["buffers_ref"].items():
            _digest_close(bufs[k].float(), d)
