"""Backward of the person-level head (first slice) on the GPU, through the C ABI.

Kernel level (same fp32 inputs on both sides => tight tolerances, 1e-4 * max|ref| unless stated):
  din_gemm_f32, din_group_layernorm_bwd_f32, din_readout_bwd_f32 vs torch autograd (fp32, GPU);
  din_dynamic_infer_bwd_f32 vs torch autograd over the oracle's restatement of dynamic_infer_ratio (CPU).
Model level: `model.train()` + cross-entropy + `backward()` on the drop-in models vs autograd over the oracle
(which tests/test_oracle_cpu.py pins against the REFERENCE model's own gradients), with identical dropout masks;
and against the committed reference-gradient fixtures directly.  The forward runs fp16 tensor-core operands
(logits within 1e-3, north_star), so model-level gradients are compared by relative L2 error per tensor, twice:
  (1) backward in isolation -- the oracle is handed the CUDA path's own feature map and fc_emb_1 output values,
      so every ReLU / arg-max decision is taken on identical numbers: ||Δ||₂ <= ISO_TOL ||ref||₂;
  (2) whole path vs the oracle's fp32 forward: fp16 rounding (1e-3) flips the ReLU mask of the ~0.1 % of
      pre-activations nearest zero, each flip moving one gradient element by its full value, so the error
      is ~ sqrt(flip fraction): ||Δ||₂ <= FULL_TOL ||ref||₂.  Measured values are printed."""
import glob
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _close(a, b, tol=1e-4, what=""):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = (a - b).abs().max().item()
    scale = max(b.abs().max().item(), 1e-12)
    assert err <= tol * scale, f"{what}: max|Δ| {err:.3e} > {tol} * max|ref| {scale:.3e}"


# ------------------------------------------------------------------------------------------------
# kernels
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,n,k", [(70, 33, 50), (128, 64, 16), (1, 27, 960), (200, 129, 7)])
def test_gemm_strided(cuda, m, n, k):
    from din_b200 import ops
    g = torch.Generator().manual_seed(m + n + k)
    a = torch.randn(m, k, generator=g).to(cuda)
    b = torch.randn(k, n, generator=g).to(cuda)
    ref = a @ b
    _close(ops.gemm_f32(a, b, m=m, n=n, k=k, a_strides=(k, 1), b_strides=(n, 1)), ref, what="NN")
    at, bt = a.t().contiguous(), b.t().contiguous()
    _close(ops.gemm_f32(at, bt, m=m, n=n, k=k, a_strides=(1, m), b_strides=(1, k)), ref, what="TT")
    bh = b.half()
    _close(ops.gemm_f32(a, bh, m=m, n=n, k=k, a_strides=(k, 1), b_strides=(n, 1)), a @ bh.float(), what="f16 B")
    out = torch.ones(m, n, device=cuda)
    ops.gemm_f32(a, b, m=m, n=n, k=k, a_strides=(k, 1), b_strides=(n, 1), out=out, alpha=0.5, accumulate=True)
    _close(out, 1 + 0.5 * ref, what="alpha/accumulate")


def test_linear_bwd(cuda):
    from din_b200 import ops
    g = torch.Generator().manual_seed(5)
    x = torch.randn(96, 48, generator=g).to(cuda).requires_grad_(True)
    w = torch.randn(20, 48, generator=g).to(cuda).requires_grad_(True)
    bias = torch.randn(20, generator=g).to(cuda).requires_grad_(True)
    dy = torch.randn(96, 20, generator=g).to(cuda)
    F.linear(x, w, bias).backward(dy)
    dx, dw, db = ops.linear_bwd(x.detach(), w.detach(), dy)
    _close(dx, x.grad, what="dx"); _close(dw, w.grad, what="dw"); _close(db, bias.grad, what="db")


LN_CASES = {
    # name: (x shape, normalized dims, pre, post, geometry kwargs builder)
    "per_person": dict(shape=(37, 64), norm=(64,), pre=False, post=False),
    "clip_wide_pre": dict(shape=(3, 4, 5, 32), norm=(4, 5, 32), pre=True, post=False),
    "clip_wide_post": dict(shape=(3, 4, 5, 32), norm=(4, 5, 32), pre=False, post=True),
}


@pytest.mark.parametrize("name", sorted(LN_CASES))
def test_group_layernorm_bwd(cuda, name):
    from din_b200 import ops
    c = LN_CASES[name]
    g = torch.Generator().manual_seed(7)
    shp, norm = c["shape"], c["norm"]
    x = torch.randn(shp, generator=g).to(cuda).requires_grad_(True)
    pre = torch.randn(shp, generator=g).to(cuda).requires_grad_(True) if c["pre"] else None
    post = torch.randn(shp, generator=g).to(cuda).requires_grad_(True) if c["post"] else None
    gamma = (torch.rand(norm, generator=g) + 0.5).to(cuda).requires_grad_(True)
    beta = (torch.randn(norm, generator=g) * 0.3).to(cuda).requires_grad_(True)
    dy = torch.randn(shp, generator=g).to(cuda)
    u = x + pre if pre is not None else x
    y = F.relu(F.layer_norm(u, norm, gamma, beta, 1e-5))
    if post is not None:
        y = y + post
    y.backward(dy)
    cols = 1
    for d in norm:
        cols *= d
    n_outer = x.numel() // cols
    dx, dg, db = ops.group_layernorm_bwd(x.detach(), gamma.detach(), beta.detach(), dy, n_outer=n_outer,
                                         outer_stride=cols, cols=cols, relu=True,
                                         pre=None if pre is None else pre.detach())
    _close(dx, x.grad, what="dx"); _close(dg, gamma.grad, what="dgamma"); _close(db, beta.grad, what="dbeta")
    if pre is not None:
        _close(dx, pre.grad, what="dpre")
    # accumulate into an existing gradient
    acc = torch.full_like(dx, 2.0)
    ops.group_layernorm_bwd(x.detach(), gamma.detach(), beta.detach(), dy, n_outer=n_outer, outer_stride=cols,
                            cols=cols, relu=True, pre=None if pre is None else pre.detach(), dx_out=acc,
                            accumulate=True, param_grads=False)
    _close(acc, x.grad + 2.0, what="dx accumulate")


def test_group_layernorm_bwd_collective_geometry(cuda):
    """LayerNorm([T, C]) per (clip, actor) over the [N, T, C] permutation of [T, N, C] with per-clip actor counts."""
    from din_b200 import ops
    B, T, N, C = 3, 4, 6, 32
    g = torch.Generator().manual_seed(9)
    x = torch.randn(B, T, N, C, generator=g).to(cuda)
    pre = torch.randn(B, T, N, C, generator=g).to(cuda)
    gamma = (torch.rand(T, C, generator=g) + 0.5).to(cuda).requires_grad_(True)
    beta = (torch.randn(T, C, generator=g) * 0.3).to(cuda).requires_grad_(True)
    dy = torch.randn(B, T, N, C, generator=g).to(cuda)
    nv = torch.tensor([6, 1, 4], dtype=torch.int32, device=cuda)
    xs = x.clone().requires_grad_(True)
    total = 0
    for b in range(B):
        n = int(nv[b])
        u = (xs[b, :, :n] + pre[b, :, :n]).permute(1, 0, 2)                         # [n, T, C]
        y = F.relu(F.layer_norm(u, (T, C), gamma, beta, 1e-5))
        total = total + (y * dy[b, :, :n].permute(1, 0, 2)).sum()
    total.backward()
    dx, dg, db = ops.group_layernorm_bwd(x, gamma.detach(), beta.detach(), dy, n_outer=B, n_inner=N,
                                         outer_stride=T * N * C, inner_stride=C, rows=T, row_stride=N * C, cols=C,
                                         relu=True, pre=pre, n_valid=nv)
    _close(dx, xs.grad, what="dx"); _close(dg, gamma.grad, what="dgamma"); _close(db, beta.grad, what="dbeta")


@pytest.mark.parametrize("with_valid", [False, True])
def test_readout_bwd(cuda, with_valid):
    from din_b200 import ops
    B, T, N, C, A = 3, 4, 5, 64, 8
    g = torch.Generator().manual_seed(3)
    s = F.relu(torch.randn(B, T, N, C, generator=g)).to(cuda).requires_grad_(True)   # zeros => arg-max ties
    w = torch.randn(A, C, generator=g).to(cuda).requires_grad_(True)
    bias = torch.randn(A, generator=g).to(cuda).requires_grad_(True)
    dl = torch.randn(B, A, generator=g).to(cuda)
    nv = torch.tensor([5, 2, 1], dtype=torch.int32, device=cuda) if with_valid else None
    rows = []
    for b in range(B):
        n = int(nv[b]) if with_valid else N
        pooled = s[b, :, :n].max(dim=1)[0]
        rows.append(F.linear(pooled, w, bias).mean(dim=0, keepdim=True))
    torch.cat(rows).backward(dl)
    ds, dw, db = ops.readout_bwd(s.detach(), w.detach(), dl, n_valid=nv)
    _close(dw, w.grad, what="dw"); _close(db, bias.grad, what="dbias")
    # ties (several actors at the pooled value, here zeros): any sub-gradient is valid; compare where unique
    sd = s.detach()
    n_lim = (nv if with_valid else torch.full((B,), N, device=cuda, dtype=torch.int32)).view(B, 1, 1, 1)
    valid = torch.arange(N, device=cuda).view(1, 1, N, 1) < n_lim
    mx = torch.where(valid, sd, torch.full_like(sd, -1e30)).max(dim=2, keepdim=True)[0]
    unique = ((sd == mx) & valid).sum(dim=2, keepdim=True) == 1
    assert unique.float().mean().item() > 0.5
    _close(torch.where(unique, ds, torch.zeros_like(ds)), torch.where(unique, s.grad, torch.zeros_like(ds)), what="ds")
    # the total gradient mass per (b,t,c) is the same with or without ties
    _close(ds.sum(dim=2), s.grad.sum(dim=2), what="ds mass")


def test_scale_mask(cuda):
    from din_b200 import ops
    x = torch.randn(1000, device=cuda)
    m = (torch.rand(1000, device=cuda) > 0.3).to(torch.uint8)
    assert torch.equal(ops.scale_mask(x, m, 1.25), x * m.float() * 1.25)
    assert torch.equal(ops.scale_mask(x, None, 2.0), x * 2.0)


DIN_CASES = [
    # kernel, ratio, scale_factor, beta(coef ptr), T, N, C, n_valid
    ((3, 3), 1, True, False, 4, 5, 32, None),
    ((3, 3), 3, True, True, 10, 12, 64, None),
    ((1, 3), 1, True, False, 4, 5, 32, None),
    ((3, 1), 2, True, False, 6, 5, 32, None),
    ((3, 3), 1, False, False, 4, 5, 32, None),
    ((3, 3), 2, True, True, 5, 13, 32, [13, 1, 4]),
]


@pytest.mark.parametrize("case", DIN_CASES, ids=[str(c[:4]) + ("_valid" if c[7] else "") for c in DIN_CASES])
def test_dynamic_infer_bwd_matches_oracle_autograd(cuda, case):
    import din_oracle as O
    from din_b200 import ops
    kernel, ratio, scale_factor, beta, T, N, C, valid = case
    kt, kn = kernel
    k2 = kt * kn
    B = 3 if valid else 2
    g = torch.Generator().manual_seed(100 + ratio + k2)
    x = torch.randn(B, T, N, C, generator=g)
    p_w = (torch.randn(2 * k2, C, kt, kn, generator=g) * 0.05).requires_grad_(True)
    p_b = (torch.randn(2 * k2, generator=g) * 1.2).requires_grad_(True)
    s_w = (torch.randn(k2, C, kt, kn, generator=g) * 0.05).requires_grad_(True) if scale_factor else None
    s_b = (torch.randn(k2, generator=g) * 0.5).requires_grad_(True) if scale_factor else None
    coef = torch.tensor([0.7]).requires_grad_(True)
    dy = torch.randn(B, T, N, C, generator=g)
    # ---- oracle: autograd over the restatement (per clip when actor counts vary, as the reference does)
    xr = x.clone().requires_grad_(True)
    total = 0
    for b in range(B):
        n = valid[b] if valid else N
        out, _ = O.din_ratio(xr[b:b + 1, :, :n], p_w, p_b, s_w, s_b, kernel, ratio)
        total = total + ((coef if beta else 0.5) * out * dy[b:b + 1, :, :n]).sum()
    total.backward()
    # ---- CUDA
    w_tap, b_cat = ops.pack_din_weights(p_w.detach().to(cuda), p_b.detach().to(cuda),
                                        None if s_w is None else s_w.detach().to(cuda),
                                        None if s_b is None else s_b.detach().to(cuda))
    nv = torch.tensor(valid, dtype=torch.int32, device=cuda) if valid else None
    xc, dyc = x.to(cuda), dy.to(cuda)
    if valid:                                     # gradient of padded actors is not defined by the reference
        for b in range(B):
            dyc[b, :, valid[b]:] = 0
    coef_d = coef.detach().to(cuda)
    dx = torch.zeros_like(xc)
    dw, db, dcoef = ops.dynamic_infer_bwd(xc, w_tap, b_cat, dyc, dx, kernel, ratio, scale_factor=scale_factor,
                                          coef=0.5, coef_ptr=coef_d.data_ptr() if beta else None, want_dcoef=beta,
                                          n_valid=nv)
    torch.cuda.synchronize()
    _close(dx, xr.grad, 2e-4, "dx")
    w_oihw = dw.permute(1, 2, 0).reshape(dw.shape[1], C, kt, kn)
    _close(w_oihw[:2 * k2], p_w.grad, 2e-4, "d p_conv.weight")
    _close(db[:2 * k2], p_b.grad, 2e-4, "d p_conv.bias")
    if scale_factor:
        _close(w_oihw[2 * k2:], s_w.grad, 2e-4, "d scale_conv.weight")
        _close(db[2 * k2:], s_b.grad, 2e-4, "d scale_conv.bias")
    if beta:
        _close(dcoef, coef.grad, 2e-4, "d beta")


# ------------------------------------------------------------------------------------------------
# whole training step
# ------------------------------------------------------------------------------------------------
def _model_and_cfg(cuda, pc, sd, dropout):
    import infer_model as IM
    from config import Config
    cfg = Config(pc.dataset)
    cfg.log_path = None
    for k in ("backbone", "image_size", "out_size", "emb_features", "num_frames", "num_boxes", "crop_size",
              "num_features_boxes", "num_activities", "lite_dim", "ST_kernel_size", "scale_factor", "beta_factor",
              "hierarchical_inference", "num_DIM"):
        setattr(cfg, k, getattr(pc, k))
    cfg.sampling_ratio = list(pc.sampling_ratio)
    cfg.num_features_gcn = pc.num_features_boxes
    cfg.train_backbone = False
    cfg.train_dropout_prob = dropout
    cls = IM.Dynamic_collective if pc.dataset == "collective" else \
        (IM.Dynamic_TCE_volleyball if pc.tce else IM.Dynamic_volleyball)
    model = cls(cfg)
    model.load_state_dict(sd, strict=True)
    model = model.to(cuda).train()
    if pc.tce:                                      # the encoder's own nn.Dropout(0.1) layers (TCE_STBiP_module.py:241):
        for m in model.multilayer_head_embfeature_context_encoding.modules():   # off, the oracle restates eval-mode TCE
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
    for m in model.modules():                       # the reference's set_bn_eval (train_net_dynamic.py:101-102)
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.eval()
    return model, cfg


def _rel_l2(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


def _pc(backbone, hw, **kw):
    import din_oracle as O
    return O.PathConfig(backbone=backbone, image_size=hw, out_size=O.backbone_out_size(backbone, *hw), **kw)


STEP_CASES = {
    "vgg16_lite": (dict(backbone="vgg16", hw=(96, 160), num_frames=3, num_boxes=4), 2, 0.3),
    "vgg16_lite_nodrop": (dict(backbone="vgg16", hw=(96, 160), num_frames=3, num_boxes=4), 2, 0.0),
    "res18_lite": (dict(backbone="res18", hw=(96, 160), num_frames=3, num_boxes=4), 2, 0.3),
    "vgg16_full_r13_beta": (dict(backbone="vgg16", hw=(96, 160), num_frames=4, num_boxes=5, lite_dim=None,
                                 sampling_ratio=(1, 3), beta_factor=True), 2, 0.3),
    "vgg16_parallel_fields": (dict(backbone="vgg16", hw=(96, 160), num_frames=4, num_boxes=5,
                                   ST_kernel_size=[(1, 3), (3, 1)], num_DIM=2), 2, 0.3),
    "vgg16_hierarchical": (dict(backbone="vgg16", hw=(64, 96), num_frames=10, num_boxes=12, lite_dim=None,
                                ST_kernel_size=[(1, 3), (3, 1)], hierarchical_inference=True), 1, 0.3),
    "collective_res18": (dict(backbone="res18", hw=(96, 144), dataset="collective", num_frames=3, num_boxes=13,
                              lite_dim=None, ST_kernel_size=(3, 3), num_activities=4), 3, 0.5),
    # Dynamic_TCE_volleyball (scripts/train_volleyball_stage2_dynamic_tce.py): context encoder in front of DIN, C = 1536
    "tce_vgg16": (dict(backbone="vgg16", hw=(96, 160), num_frames=3, num_boxes=12, lite_dim=None, tce=True), 2, 0.3),
}


@pytest.mark.parametrize("name", sorted(STEP_CASES))
def test_training_step_matches_oracle(cuda, name):
    """model.train(); loss = cross_entropy(model(batch)); loss.backward()  vs autograd over the oracle, with the
    same dropout masks (drawn from torch's CUDA generator in the order the path draws them)."""
    import din_oracle as O
    from din_b200 import metrics
    kw, B, p = STEP_CASES[name]
    kw = dict(kw)
    pc = _pc(kw.pop("backbone"), kw.pop("hw"), **kw)
    bb = O.build_backbone(pc.backbone)
    seed = int(os.environ.get("DIN_TEST_SEED", "0"))         # (tolerance studies: the same test over other weights / inputs)
    sd = O.make_state_dict(pc, seed=seed, backbone=bb)
    O.load_backbone(bb, sd)
    bb.eval()
    batch = O.make_inputs(pc, B, seed=seed)
    labels = torch.arange(B) % pc.num_activities
    model, cfg = _model_and_cfg(cuda, pc, sd, p)
    model.keep_tape = True
    T, N, C = pc.num_frames, pc.num_boxes, pc.in_dim
    # ---- CUDA step
    torch.manual_seed(1234)
    out = model(tuple(t.to(cuda) for t in batch))["activities"]
    loss = metrics.cross_entropy(out, labels.to(cuda))
    loss.backward()
    torch.cuda.synchronize()
    # ---- the same masks for the oracle
    torch.manual_seed(1234)
    hmask = None
    if pc.hierarchical_inference:
        hmask = (torch.rand((B, T, N, C), device=cuda) >= 0.5).cpu()
    mask = (torch.rand((B, T, N, C), device=cuda) >= p).cpu() if p > 0 else None
    train = {"p": p, "mask": mask, "hmask": hmask}
    got = {n: q.grad for n, q in model.named_parameters() if q.grad is not None}
    assert all(q.grad is None for n, q in model.named_parameters() if n.startswith("backbone."))

    # (1) the backward in isolation: the oracle is handed the CUDA path's own feature map and the VALUE of its
    #     fc_emb_1 output (the only fp16 tensor-core product after the backbone), so that every ReLU / arg-max
    #     decision downstream is taken on the same numbers; what is left is fp32 summation order and the fp16
    #     rounding of the crops inside d(fc_emb_1.weight)
    fm = model.engine()._fm_cache[1][..., :pc.emb_features].float().permute(0, 3, 1, 2).contiguous().cpu()
    iso_logits, iso_loss, iso_grads = O.head_grads(bb, sd, pc, labels, *batch, features=fm,
                                                   train=dict(train, emb=model._last_tape["emb"].cpu()))
    assert set(got) == set(iso_grads), set(got) ^ set(iso_grads)
    worst_iso = max(_rel_l2(got[k], iso_grads[k]) for k in iso_grads)
    for k in sorted(iso_grads):
        print(f"[step {name}] {k:45s} isolated rel-L2 {_rel_l2(got[k], iso_grads[k]):.2e}")
    err_iso = (out.detach().cpu() - iso_logits).abs().max().item()
    print(f"[step {name}] isolated: logits max|Δ| {err_iso:.2e}, worst rel-L2 {worst_iso:.2e}")

    # (2) the whole path against the oracle's own fp32 backbone: adds the fp16 tensor-core backbone's rounding,
    #     which also flips a few ReLU / arg-max decisions (measured values printed)
    ref_logits, ref_loss, ref_grads = O.head_grads(bb, sd, pc, labels, *batch, train=train)
    err = (out.detach().cpu() - ref_logits).abs().max().item()
    worst = max(_rel_l2(got[k], ref_grads[k]) for k in ref_grads)
    for k in sorted(ref_grads):
        print(f"[step {name}] {k:45s} rel-L2 {_rel_l2(got[k], ref_grads[k]):.2e}  |ref| {float(ref_grads[k].norm()):.3e}")
    print(f"[step {name}] logits max|Δ| {err:.2e}, loss {loss.item():.6f} vs {ref_loss.item():.6f}, worst rel-L2 {worst:.2e}")
    assert err_iso <= 1e-3 * iso_logits.abs().max().item(), ("isolated logits", err_iso)
    # TCE: the heads' downsample2 GEMM (fp16 tensor-core operands) sits between the shared feature map and the softmax of
    # the attention, which amplifies its rounding: the encoder's gradients agree to a few 1e-2 even "in isolation", and how
    # few depends on the realisation -- over DIN_TEST_SEED = 0..4 the worst tensor's relative L2 error was 1.3e-2, 3.8e-2,
    # 2.3e-2, 2.5e-2, 1.9e-1 with the convolutions' bias in the epilogue and 3.6e-2, 3.4e-2, 3.6e-2, 5.8e-2, 2.6e-2 with
    # it on the tensor core (profiles/edge_precision_study_r2.md); seed 0 is what runs here
    assert worst_iso <= (6e-2 if pc.tce else ISO_TOL), ("isolated", worst_iso)
    assert err <= 1e-3 * ref_logits.abs().max().item(), ("logits", err)
    assert abs(loss.item() - ref_loss.item()) <= 2e-3 * max(1.0, abs(ref_loss.item()))
    # TCE: the attention softmax amplifies the fp16 backbone's rounding (emb_roi / the DIN offset convolutions: 1.1e-1)
    assert worst <= (1.5e-1 if pc.tce else FULL_TOL), ("whole path", worst)


ISO_TOL = 3e-3      # relative L2 per gradient tensor, backward in isolation
FULL_TOL = 8e-2     # relative L2 per gradient tensor, fp16 forward included (ReLU-mask flips, see the docstring)


GRAD_FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "grads_*.pt")))


@pytest.mark.parametrize("path", GRAD_FIXTURES, ids=[os.path.basename(p) for p in GRAD_FIXTURES])
def test_training_step_matches_reference_fixture(cuda, path):
    """CUDA gradients vs the gradients the REFERENCE model produced (tests/golden/grads_*.pt; dropout 0)."""
    import din_oracle as O
    from din_b200 import metrics
    from test_oracle_cpu import _pc_from
    fx = torch.load(path)
    pc = _pc_from(fx["config"])
    sd = O.make_state_dict(pc, seed=fx["seed"])
    batch = O.make_inputs(pc, fx["B"], seed=fx["seed"])
    model, _ = _model_and_cfg(cuda, pc, sd, 0.0)
    out = model(tuple(t.to(cuda) for t in batch))["activities"]
    loss = metrics.cross_entropy(out, fx["labels"].to(cuda))
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - float(fx["loss_ref"])) <= 2e-3 * max(1.0, abs(float(fx["loss_ref"])))
    got = {n: q.grad for n, q in model.named_parameters() if q.grad is not None}
    assert set(got) == set(fx["grads_ref"])
    for k, d in fx["grads_ref"].items():
        flat = got[k].detach().double().cpu().flatten()
        assert tuple(got[k].shape) == tuple(d["shape"])
        l2 = float(flat.norm())
        smp = (flat[d["idx"]].float() - d["samples"]).abs().max().item()
        print(f"[fixture] {k:45s} l2 {l2:.4e} vs {d['l2']:.4e}   samples max|Δ| {smp:.2e} (max|ref| {d['max_abs']:.2e})")
        assert abs(l2 - d["l2"]) <= FULL_TOL * max(d["l2"], 1e-30), (k, l2, d["l2"])
        # a flipped ReLU decision moves a single element by its full value: require 85 % of the sampled elements
        # within 2e-2 * max|ref| rather than all of them (the smallest tensors have 9 elements)
        near = ((flat[d["idx"]].float() - d["samples"]).abs() <= 2e-2 * max(d["max_abs"], 1e-30)).float().mean().item()
        # the smallest tensors (the 9 / 18 offset and relation biases) have 9 / 18 samples: two thirds is the bar there --
        # which of those few elements a flipped decision lands on changes with any change of the forward's rounding
        # (profiles/edge_precision_study_r2.md; the space-to-depth ResNet stem moved DPI.p_conv.1.bias of the Collective
        # fixture from 14 to 13 of 18), the tensor's norm above does not
        assert near >= (0.66 if len(d["idx"]) <= 32 else 0.85), (k, near, smp, d["max_abs"])


def test_optimizer_loop_and_modes(cuda):
    """Two Adam steps as train_net_dynamic.py:170-224 runs them (weights change => the plan is rebuilt), then
    eval; training the backbone or batch-stat BatchNorm raise instead of silently doing something else."""
    import din_oracle as O
    from din_b200 import metrics
    pc = _pc("vgg16", (96, 160), num_frames=3, num_boxes=4)
    sd = O.make_state_dict(pc, seed=0)
    batch = tuple(t.to(cuda) for t in O.make_inputs(pc, 2, seed=0))
    labels = torch.tensor([1, 5], device=cuda)
    model, cfg = _model_and_cfg(cuda, pc, sd, 0.3)
    params = [q for q in model.parameters() if q.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-3)
    meters = metrics.DeviceMeters(pc.num_activities, cuda)
    losses = []
    before = model.fc_activities.weight.detach().clone()
    for _ in range(3):
        out = model(batch)["activities"]
        loss = metrics.cross_entropy(out, labels, meters=meters)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(torch.isfinite(torch.tensor(losses))) and not torch.equal(before, model.fc_activities.weight)
    assert meters.value()["steps"] == 3
    model.eval()
    with torch.no_grad():
        assert torch.isfinite(model(batch)["activities"]).all()
    # Inception-v3 with batch-statistics BatchNorm is not implemented: a loud error
    pc3 = _pc("inv3", (139, 203), emb_features=1056, num_frames=2, num_boxes=4, lite_dim=None)
    m3, _ = _model_and_cfg(cuda, pc3, O.make_state_dict(pc3, seed=0), 0.3)
    for q in m3.backbone.parameters():
        q.requires_grad = True
    m3.train()                                        # BatchNorm back to batch statistics: ResNet-18 only
    with pytest.raises(NotImplementedError, match="BatchNorm"):
        m3(tuple(t.to(cuda) for t in O.make_inputs(pc3, 2, seed=0)))


@pytest.mark.parametrize("case", ["vgg16_lite", "res18_lite", "collective_res18", "inv3_full"])
def test_full_training_step_with_backbone(cuda, case):
    """cfg.train_backbone = True (scripts/train_volleyball_stage2_dynamic.py:12) on VGG-16 (43 parameter tensors),
    ResNet-18 (77, BatchNorm in eval mode: gamma / beta still train) and Inception-v3 (223: 94 folded conv + BatchNorm
    pairs, multiscale map): every gradient vs autograd over the oracle, and vs the REFERENCE model's gradients (fixture).
    The backbone's backward runs on fp16 tensor-core operands (dynamic loss scale): relative L2 per tensor."""
    import din_oracle as O
    from din_b200 import metrics
    from test_oracle_cpu import _pc_from
    fx = torch.load(os.path.join(GOLDEN, f"fullgrads_{case}.pt"))
    pc = _pc_from(fx["config"])
    bb = O.build_backbone(pc.backbone)
    sd = O.make_state_dict(pc, seed=fx["seed"], backbone=bb)
    O.load_backbone(bb, sd)
    bb.eval()
    batch = O.make_inputs(pc, fx["B"], seed=fx["seed"])
    labels = fx["labels"]
    model, cfg = _model_and_cfg(cuda, pc, sd, 0.0)
    for q in model.backbone.parameters():
        q.requires_grad = True
    out = model(tuple(t.to(cuda) for t in batch))["activities"]
    loss = metrics.cross_entropy(out, labels.to(cuda))
    loss.backward()
    torch.cuda.synchronize()
    got = {n: q.grad for n, q in model.named_parameters() if q.grad is not None}
    ref_logits, ref_loss, ref_grads = O.head_grads(bb, sd, pc, labels, *batch, train_backbone=True)
    assert set(got) == set(ref_grads) == set(fx["grads_ref"]), set(got) ^ set(ref_grads)
    worst = worst_norm = 0.0
    for k in sorted(ref_grads):
        r = _rel_l2(got[k], ref_grads[k])
        l2 = float(got[k].double().norm())
        worst = max(worst, r)
        print(f"[full step] {k:45s} rel-L2 {r:.2e}  |ref| {float(ref_grads[k].norm()):.3e}  fixture l2 {fx['grads_ref'][k]['l2']:.3e}")
        norm_err = abs(l2 - fx["grads_ref"][k]["l2"]) / fx["grads_ref"][k]["l2"]
        worst_norm = max(worst_norm, norm_err)
    print(f"[full step] loss {loss.item():.6f} vs {ref_loss.item():.6f}, worst rel-L2 {worst:.2e}, "
          f"worst norm error vs the reference fixture {worst_norm:.2e}")
    assert abs(loss.item() - ref_loss.item()) <= 2e-3 * max(1.0, abs(ref_loss.item()))
    assert worst <= BB_TOL, worst
    # every tensor's norm within 5 % of the reference's (measured: VGG-16 1.4e-2, ResNet-18 1.7e-2, Collective ResNet-18
    # 2.6e-2 .. 3.04e-2 over runs -- fp32 atomics make the last digits vary)
    # Inception-v3 (94 convolutions, 223 tensors, the deepest chain 37 layers): 6.1e-2 on its worst tensor (a BatchNorm
    # weight near the stem), rel-L2 growing from 2e-3 at the head to 1.6e-1 at Conv2d_1a like VGG-16's 1.3e-1 at features.0
    assert worst_norm <= (8e-2 if case == "inv3_full" else 5e-2), worst_norm
    # one optimizer step over everything, then a second forward (weights repacked)
    opt = torch.optim.Adam([q for q in model.parameters() if q.requires_grad], lr=1e-4)
    opt.step()
    out2 = model(tuple(t.to(cuda) for t in batch))["activities"]
    assert torch.isfinite(out2).all() and not torch.equal(out2, out)


# Relative L2 per gradient tensor with the backbone trained.  Every layer's fp16 forward flips the ReLU / max-pool
# decision of the ~0.1 % of units nearest a tie, and a flipped unit reroutes its whole gradient, so the error grows
# like sqrt(#layers passed): measured 2.6e-2 at features.28 (last conv), 3.7e-2 at features.19, 7e-2 at
# features.7, 1.3e-1 at features.0 -- while every tensor's NORM agrees with the reference's to 3-4 digits and each
# kernel on the way is exact or 1e-4-tight in isolation (tests/test_conv_bwd_gpu.py).
BB_TOL = 2e-1


@pytest.mark.parametrize("hw,T,N", [((96, 160), 3, 4), ((720, 1280), 2, 12)], ids=["small", "720p"])
def test_gradient_of_batch_is_mean_of_per_clip_gradients(cuda, hw, T, N):
    """Size-independent property behind the data-parallel step, also at BASELINE's frame size (where
    the CPU oracle would take minutes): with dropout off, the gradient of the batch-mean loss equals the mean of the
    per-clip gradients (what the single flat all-reduce averages), backbone included; and two identical steps agree
    to atomics noise."""
    import din_oracle as O
    from din_b200 import metrics
    pc = _pc("vgg16", hw, num_frames=T, num_boxes=N)
    sd = O.make_state_dict(pc, seed=4)
    batch = tuple(t.to(cuda) for t in O.make_inputs(pc, 2, seed=4))
    labels = torch.tensor([3, 6], device=cuda)
    model, _ = _model_and_cfg(cuda, pc, sd, 0.0)
    for q in model.backbone.parameters():
        q.requires_grad = True

    def grads(clips, lab):
        for q in model.parameters():
            q.grad = None
        metrics.cross_entropy(model(clips)["activities"], lab).backward()
        return {n: q.grad.clone() for n, q in model.named_parameters()}

    full = grads(batch, labels)
    again = grads(batch, labels)
    g0 = grads(tuple(t[:1] for t in batch), labels[:1])
    g1 = grads(tuple(t[1:] for t in batch), labels[1:])
    for k in full:
        # fp32 atomics (RoIAlign / walk scatter, split-K wgrad) reorder sums; the fp16 rounding of the backbone
        # gradients then amplifies the last-bit differences: 8.5e-4 measured at features.0
        assert _rel_l2(again[k], full[k]) <= 5e-3, ("run-to-run", k, _rel_l2(again[k], full[k]))
        # the dynamic loss scale differs between the runs (per-step max|dfm|), hence fp16 rounding differs: 2e-3
        assert _rel_l2(0.5 * (g0[k] + g1[k]), full[k]) <= 1e-2, (k, _rel_l2(0.5 * (g0[k] + g1[k]), full[k]))


def test_training_reduces_loss_on_a_fixed_batch(cuda):
    """End-to-end sanity of the gradients' SIGN and scale: 12 Adam steps on one fixed batch (everything trained, dropout
    on, as train_net_dynamic.py:170-224 runs them) must overfit it."""
    import din_oracle as O
    from din_b200 import metrics
    pc = _pc("vgg16", (96, 160), num_frames=3, num_boxes=4)
    sd = O.make_state_dict(pc, seed=5)
    batch = tuple(t.to(cuda) for t in O.make_inputs(pc, 4, seed=5))
    labels = torch.tensor([0, 3, 5, 7], device=cuda)
    model, _ = _model_and_cfg(cuda, pc, sd, 0.3)
    for q in model.backbone.parameters():
        q.requires_grad = True
    opt = torch.optim.Adam([q for q in model.parameters() if q.requires_grad], lr=2e-4)
    torch.manual_seed(0)
    losses = []
    for _ in range(12):
        loss = metrics.cross_entropy(model(batch)["activities"], labels)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    print(f"\n[overfit] loss {losses[0]:.3f} -> {losses[-1]:.3f}  ({[round(v, 2) for v in losses]})")
    assert all(torch.isfinite(torch.tensor(losses)))
    assert min(losses[-3:]) < 0.5 * losses[0], losses


def test_collective_training_step_with_vgg16_backbone(cuda):
    """Dynamic_collective (variable actor counts, LayerNorm([T, C]) per actor) with cfg.train_backbone = True on the
    VGG-16 backbone: every gradient vs autograd over the oracle (fp16 backbone tolerance, see BB_TOL)."""
    import din_oracle as O
    from din_b200 import metrics
    pc = _pc("vgg16", (96, 144), dataset="collective", num_frames=3, num_boxes=13, lite_dim=None, ST_kernel_size=(3, 3),
             num_activities=4)
    bb = O.build_backbone(pc.backbone)
    sd = O.make_state_dict(pc, seed=6, backbone=bb)
    O.load_backbone(bb, sd)
    bb.eval()
    batch = O.make_inputs(pc, 3, seed=6)
    labels = torch.tensor([0, 3, 1])
    model, _ = _model_and_cfg(cuda, pc, sd, 0.0)
    for q in model.backbone.parameters():
        q.requires_grad = True
    out = model(tuple(t.to(cuda) for t in batch))["activities"]
    loss = metrics.cross_entropy(out, labels.to(cuda))
    loss.backward()
    torch.cuda.synchronize()
    ref_logits, ref_loss, ref_grads = O.head_grads(bb, sd, pc, labels, *batch, train_backbone=True)
    got = {n: q.grad for n, q in model.named_parameters() if q.grad is not None}
    assert set(got) == set(ref_grads), set(got) ^ set(ref_grads)
    worst = max(_rel_l2(got[k], ref_grads[k]) for k in ref_grads)
    worst_norm = max(abs(float(got[k].double().norm()) - float(ref_grads[k].double().norm())) /
                     max(float(ref_grads[k].double().norm()), 1e-30) for k in ref_grads)
    print(f"\n[collective full step] loss {loss.item():.5f} vs {ref_loss.item():.5f}; worst rel-L2 {worst:.2e}; worst "
          f"norm error {worst_norm:.2e}")
    assert abs(loss.item() - ref_loss.item()) <= 2e-3 * max(1.0, abs(ref_loss.item()))
    assert worst <= BB_TOL and worst_norm <= 5e-2, (worst, worst_norm)


@pytest.mark.parametrize("case", ["res18_lite", "collective_res18"])
def test_training_step_with_batchnorm_on_batch_statistics(cuda, case):
    """ResNet-18 trained WITHOUT cfg.set_bn_eval (config.py:80 default; scripts/train_collective_stage2_dynamic.py):
    every BatchNorm normalises with the statistics of the step's B*T frames and updates its running statistics.
    Gradients vs autograd over the oracle (bb.train()) and vs the reference model's own (fixture); running statistics vs
    the oracle's; eval() afterwards folds the UPDATED statistics."""
    import din_oracle as O
    from din_b200 import metrics
    from test_oracle_cpu import _pc_from
    fx = torch.load(os.path.join(GOLDEN, f"bntrain_{case}.pt"))
    pc = _pc_from(fx["config"])
    bb = O.build_backbone(pc.backbone)
    sd = O.make_state_dict(pc, seed=fx["seed"], backbone=bb)
    O.load_backbone(bb, sd)
    bb.train()
    batch = O.make_inputs(pc, fx["B"], seed=fx["seed"])
    labels = fx["labels"]
    model, cfg = _model_and_cfg(cuda, pc, sd, 0.0)
    model.train()                                                    # BatchNorm layers back on batch statistics
    for q in model.backbone.parameters():
        q.requires_grad = True
    gpu_batch = tuple(t.to(cuda) for t in batch)
    with torch.no_grad():
        eval_before = model.eval()(gpu_batch)["activities"].clone()
    model.train()
    out = model(gpu_batch)["activities"]
    loss = metrics.cross_entropy(out, labels.to(cuda))
    loss.backward()
    torch.cuda.synchronize()
    got = {n: q.grad for n, q in model.named_parameters() if q.grad is not None}
    ref_logits, ref_loss, ref_grads = O.head_grads(bb, sd, pc, labels, *batch, train_backbone=True)
    assert set(got) == set(ref_grads) == set(fx["grads_ref"]), set(got) ^ set(ref_grads)
    print(f"[bn-train {case}] logits max|d| {(out.detach().cpu() - ref_logits).abs().max().item():.2e}, "
          f"loss {loss.item():.6f} vs {ref_loss.item():.6f} (reference fixture {float(fx['loss_ref']):.6f})")
    assert abs(loss.item() - ref_loss.item()) <= 3e-3 * max(1.0, abs(ref_loss.item()))
    worst = worst_norm = 0.0
    for k in sorted(ref_grads):
        r = _rel_l2(got[k], ref_grads[k])
        l2 = float(got[k].double().norm())
        norm_err = abs(l2 - fx["grads_ref"][k]["l2"]) / fx["grads_ref"][k]["l2"]
        worst, worst_norm = max(worst, r), max(worst_norm, norm_err)
        print(f"[bn-train {case}] {k:45s} rel-L2 {r:.2e}  norm vs reference {norm_err:.2e}")
    print(f"[bn-train {case}] worst rel-L2 {worst:.2e}, worst norm error vs the reference fixture {worst_norm:.2e}")
    # batch statistics make the network ~3x more sensitive to fp16-sized perturbations than eval-mode BatchNorm: the
    # ORACLE's own gradients move by 1.9e-1 (worst tensor; eval mode 6.7e-2) when every conv weight is perturbed by
    # 3e-4 relative (tests/tools/bn_sensitivity.py); measured here 1.8e-1 .. 1.9e-1, norms within 5.6e-2 of the reference's
    assert worst <= 3e-1, worst
    assert worst_norm <= 8e-2, worst_norm
    # running statistics: one momentum step from (0, 1) towards the batch statistics
    bufs = dict(model.named_buffers())
    n_bn = 0
    for n, b in bb.named_buffers():
        mine = bufs["backbone." + n].detach().cpu()
        if n.endswith("num_batches_tracked"):
            assert int(mine) == int(b) == 1, (n, int(mine), int(b))   # one train-mode forward each; eval does not count
            continue
        n_bn += 1
        assert torch.allclose(mine, b, rtol=5e-3, atol=2e-4), (n, (mine - b).abs().max().item())
    assert n_bn == 40
    # eval() now folds the updated statistics: the logits move
    with torch.no_grad():
        eval_after = model.eval()(gpu_batch)["activities"]
    assert torch.isfinite(eval_after).all() and not torch.allclose(eval_after, eval_before)
    # frozen backbone, BatchNorm still on batch statistics (scripts/train_volleyball_stage2_arg.py: train_backbone False)
    model.train()
    for q in model.backbone.parameters():
        q.requires_grad = False
        q.grad = None
    out2 = model(gpu_batch)["activities"]
    out2.sum().backward()
    assert torch.isfinite(out2).all() and all(q.grad is None for q in model.backbone.parameters())


def test_tce_training_step_with_backbone_and_encoder_dropout(cuda):
    """Dynamic_TCE_volleyball with the VGG-16 backbone trained: the context encoder sends its own share of the
    feature-map gradient (through the heads' downsample2 GEMM) to the backbone; every gradient vs autograd over the
    oracle.  Then one step with the encoder's dropout layers active (p = 0.1): finite, and different."""
    import din_oracle as O
    from din_b200 import metrics
    pc = _pc("vgg16", (96, 160), num_frames=3, num_boxes=12, lite_dim=None, tce=True)
    bb = O.build_backbone(pc.backbone)
    sd = O.make_state_dict(pc, seed=0, backbone=bb)
    O.load_backbone(bb, sd)
    bb.eval()
    B = 2
    batch = O.make_inputs(pc, B, seed=0)
    labels = torch.arange(B) % pc.num_activities
    model, cfg = _model_and_cfg(cuda, pc, sd, 0.0)
    for q in model.backbone.parameters():
        q.requires_grad = True
    gpu_batch = tuple(t.to(cuda) for t in batch)
    out = model(gpu_batch)["activities"]
    loss = metrics.cross_entropy(out, labels.to(cuda))
    loss.backward()
    torch.cuda.synchronize()
    got = {n: q.grad.clone() for n, q in model.named_parameters() if q.grad is not None}
    ref_logits, ref_loss, ref_grads = O.head_grads(bb, sd, pc, labels, *batch, train_backbone=True)
    assert set(got) == set(ref_grads), set(got) ^ set(ref_grads)
    worst = 0.0
    for k in sorted(ref_grads):
        r = _rel_l2(got[k], ref_grads[k])
        worst = max(worst, r)
        if "context_encoding" in k or r > 5e-2:
            print(f"[tce full step] {k:70s} rel-L2 {r:.2e}")
    print(f"[tce full step] loss {loss.item():.6f} vs {ref_loss.item():.6f}, worst rel-L2 {worst:.2e}")
    assert abs(loss.item() - ref_loss.item()) <= 2e-3 * max(1.0, abs(ref_loss.item()))
    assert worst <= BB_TOL, worst
    # the encoder's dropout layers back on
    for m in model.multilayer_head_embfeature_context_encoding.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.1
    model.zero_grad()
    torch.manual_seed(7)
    loss2 = metrics.cross_entropy(model(gpu_batch)["activities"], labels.to(cuda))
    loss2.backward()
    torch.cuda.synchronize()
    assert torch.isfinite(loss2) and abs(loss2.item() - loss.item()) > 1e-6
    assert all(torch.isfinite(q.grad).all() for q in model.parameters() if q.grad is not None)
