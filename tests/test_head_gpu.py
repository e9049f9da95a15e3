"""Person-level head kernels (RoIAlign, LayerNorms, linear, fused Dynamic Relation / Walk, read-out) vs
the CPU oracle (oracle/din_oracle.py) on identical seeded inputs.  fp32 kernels: tolerance 2e-5 relative
to max|ref| (summation-order differences only); RoIAlign reads an fp16 map, and the test hands the oracle
the same fp16-rounded values, output rounded to fp16: tolerance 1e-3 * max|ref|."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _close(a, ref, rel, what=""):
    err = (a.detach().cpu().float() - ref).abs().max().item()
    scale = max(ref.abs().max().item(), 1e-6)
    assert err <= rel * scale, f"{what}: max abs err {err:.3e} vs max|ref| {scale:.3e} (rel tol {rel})"


def test_roi_align_matches_oracle(cuda):
    import din_oracle as O
    from din_b200 import ops
    g = torch.Generator().manual_seed(5)
    n_img, H, W, D, M = 3, 22, 40, 64, 60
    fm = torch.randn(n_img, D, H, W, generator=g).half().float()
    cx, cy = torch.rand(M, generator=g) * W, torch.rand(M, generator=g) * H
    bw, bh = 1 + 3 * torch.rand(M, generator=g), 2 + 5 * torch.rand(M, generator=g)
    boxes = torch.stack((cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2), 1)
    boxes[0] = torch.tensor([0.0, 0.0, 0.0, 0.0])             # Collective's padded box
    boxes[1] = torch.tensor([-3.0, -2.0, 4.0, 5.0])           # straddles the top-left border
    boxes[2] = torch.tensor([W - 2.0, H - 3.0, W + 3.0, H + 2.0])
    boxes[3] = torch.tensor([2.0, 3.0, 7.0, 8.0])             # integer-aligned sample points
    idx = torch.randint(0, n_img, (M,), generator=g).int()
    ref = O.roi_align_longcw(fm, boxes, idx, 5, 5)            # [M,D,5,5]
    out = ops.roi_align_nhwc(fm.permute(0, 2, 3, 1).contiguous().half().to(cuda), boxes.to(cuda), idx.to(cuda), 5, 5)
    got = out.view(M, 5, 5, D).permute(0, 3, 1, 2)
    _close(got, ref, 1e-3, "roi_align")
    # zero pattern (extrapolation) must agree exactly
    assert torch.equal(got.cpu() == 0, ref.half() == 0)
    # the stand-alone drop-in module (NCHW fp32 in/out) wraps the same kernel
    from roi_align.roi_align import RoIAlign
    got2 = RoIAlign(5, 5)(fm.to(cuda), boxes.to(cuda), idx.to(cuda))
    _close(got2, ref, 1e-3, "RoIAlign module")


def test_roi_align_vs_torchvision_interior(cuda):
    """Independent cross-check of the (unpinned) RoIAlign restatement: on boxes whose sample points stay
    inside the map, crop_and_resize == torchvision roi_align(aligned=True, sampling_ratio=1)."""
    import din_oracle as O
    from torchvision.ops import roi_align as tv_roi_align
    g = torch.Generator().manual_seed(6)
    n_img, H, W, D, M = 2, 30, 44, 16, 40
    fm = torch.randn(n_img, D, H, W, generator=g)
    cx, cy = 8 + torch.rand(M, generator=g) * (W - 16), 8 + torch.rand(M, generator=g) * (H - 16)
    bw, bh = 1 + 3 * torch.rand(M, generator=g), 2 + 5 * torch.rand(M, generator=g)
    boxes = torch.stack((cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2), 1)
    idx = torch.randint(0, n_img, (M,), generator=g).int()
    ref = O.roi_align_longcw(fm, boxes, idx, 5, 5)
    rois = torch.cat((idx.float().unsqueeze(1), boxes), 1)
    tv = tv_roi_align(fm, rois, (5, 5), spatial_scale=1.0, sampling_ratio=1, aligned=True)
    assert (ref - tv).abs().max().item() < 1e-4


def test_layernorm_and_linear(cuda):
    from din_b200 import ops
    g = torch.Generator().manual_seed(8)
    # row LN + ReLU (nl_emb_1)
    x = torch.randn(37, 1024, generator=g) * 3 + 1
    w, b = torch.rand(1024, generator=g) + 0.5, torch.randn(1024, generator=g)
    ref = F.relu(F.layer_norm(x, (1024,), w, b))
    got = ops.group_layernorm(x.to(cuda), w.to(cuda), b.to(cuda), n_outer=37, outer_stride=1024, cols=1024, relu=True)
    _close(got, ref, 2e-5, "row LN")
    # clip-wide LN with pre-add (vgg16 fusion) and with post-add (res18 fusion)
    B, T, N, C = 3, 4, 5, 128
    xg, xr = torch.randn(B, T, N, C, generator=g), torch.randn(B, T, N, C, generator=g)
    w, b = torch.rand(T, N, C, generator=g) + 0.5, torch.randn(T, N, C, generator=g)
    gs = T * N * C
    ref = F.relu(F.layer_norm(xg + xr, (T, N, C), w, b))
    got = ops.group_layernorm(xg.to(cuda), w.to(cuda), b.to(cuda), n_outer=B, outer_stride=gs, cols=gs, relu=True,
                              pre=xr.to(cuda))
    _close(got, ref, 2e-5, "clip LN pre")
    ref = F.relu(F.layer_norm(xg, (T, N, C), w, b)) + xr
    got = ops.group_layernorm(xg.to(cuda), w.to(cuda), b.to(cuda), n_outer=B, outer_stride=gs, cols=gs, relu=True,
                              post=xr.to(cuda))
    _close(got, ref, 2e-5, "clip LN post")
    # Collective: LayerNorm([T, C]) over the [N, T, C] permutation, only the first n_valid actors
    w2, b2 = torch.rand(T, C, generator=g) + 0.5, torch.randn(T, C, generator=g)
    nv = torch.tensor([5, 2, 1], dtype=torch.int32)
    out = torch.zeros(B, T, N, C, device=cuda)
    ops.group_layernorm(xg.to(cuda), w2.to(cuda), b2.to(cuda), n_outer=B, n_inner=N, outer_stride=gs,
                        inner_stride=C, rows=T, row_stride=N * C, cols=C, relu=True, pre=xr.to(cuda),
                        n_valid=nv.to(cuda), out=out)
    for bi in range(B):
        n = int(nv[bi])
        s = (xg[bi, :, :n] + xr[bi, :, :n]).permute(1, 0, 2)                     # [n,T,C]
        ref = F.relu(F.layer_norm(s, (T, C), w2, b2)).permute(1, 0, 2)            # back to [T,n,C]
        _close(out[bi, :, :n], ref, 2e-5, f"collective LN clip {bi}")
        assert (out[bi, :, n:] == 0).all()
    # linear (+bias, accumulate)
    x = torch.randn(130, 1024, generator=g)
    wl, bl = torch.randn(128, 1024, generator=g) * 0.05, torch.randn(128, generator=g)
    got = ops.linear_f32(x.to(cuda), wl.to(cuda), bl.to(cuda))
    _close(got, F.linear(x, wl, bl), 2e-5, "linear")
    got = ops.linear_f32(x.to(cuda), wl.to(cuda), None, out=got, accumulate=True)
    _close(got, F.linear(x, wl, bl) + F.linear(x, wl), 2e-5, "linear accumulate")
    x3 = torch.randn(7, 48, generator=g)
    w3 = torch.randn(8, 48, generator=g)
    _close(ops.linear_f32(x3.to(cuda), w3.to(cuda), None, relu=True), F.relu(F.linear(x3, w3)), 2e-5, "linear small")


DIN_CASES = [
    # B, T, N, C, kernel, ratio, scale_factor, offset_bias_std
    (2, 10, 12, 128, (3, 3), 1, True, 0.3),
    (2, 10, 12, 128, (3, 3), 3, True, 2.5),     # large offsets: every clamp / border double-count exercised
    (1, 10, 12, 1024, (3, 3), 1, True, 1.0),
    (2, 4, 5, 64, (1, 3), 1, True, 1.5),        # ST-factorised: no padding along T -> border is real data
    (2, 4, 5, 64, (3, 1), 2, True, 1.5),
    (2, 3, 13, 64, (3, 3), 1, False, 1.0),      # scale_factor=False -> plain mean over taps
]


@pytest.mark.parametrize("case", DIN_CASES, ids=[str(c) for c in DIN_CASES])
def test_dynamic_infer_matches_oracle(cuda, case):
    import din_oracle as O
    from din_b200 import ops
    B, T, N, C, kernel, ratio, sf, ostd = case
    kt, kn = kernel
    k2 = kt * kn
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, T, N, C, generator=g)
    p_w = torch.randn(2 * k2, C, kt, kn, generator=g) * 0.01
    p_b = torch.randn(2 * k2, generator=g) * ostd
    s_w = torch.randn(k2, C, kt, kn, generator=g) * 0.02 if sf else None
    s_b = torch.randn(k2, generator=g) * 0.3 if sf else None
    ref, _ = O.din_ratio(x, p_w, p_b, s_w, s_b, kernel, ratio)
    w_tap, b_cat = ops.pack_din_weights(p_w.to(cuda), p_b.to(cuda), s_w.to(cuda) if sf else None,
                                        s_b.to(cuda) if sf else None)
    got = ops.dynamic_infer(x.to(cuda), w_tap, b_cat, kernel, ratio, scale_factor=sf)
    _close(got, ref, 2e-5, "dynamic_infer")
    # accumulate with a coefficient read from device memory (beta)
    beta = torch.tensor([0.25, 1.75], device=cuda)
    got2 = ops.dynamic_infer(x.to(cuda), w_tap, b_cat, kernel, ratio, scale_factor=sf, out=got.clone(),
                             coef_ptr=beta.data_ptr() + 4, accumulate=True)
    _close(got2, ref * 2.75, 2e-5, "dynamic_infer accumulate")


def test_dynamic_infer_variable_actors(cuda):
    """Collective: per-clip actor counts handled in one launch == per-clip calls on the sliced tensor."""
    import din_oracle as O
    from din_b200 import ops
    g = torch.Generator().manual_seed(12)
    B, T, N, C = 4, 3, 13, 64
    x = torch.randn(B, T, N, C, generator=g)
    p_w, p_b = torch.randn(18, C, 3, 3, generator=g) * 0.01, torch.randn(18, generator=g) * 1.5
    s_w, s_b = torch.randn(9, C, 3, 3, generator=g) * 0.02, torch.randn(9, generator=g) * 0.3
    nv = torch.tensor([13, 1, 4, 7], dtype=torch.int32)
    w_tap, b_cat = ops.pack_din_weights(p_w.to(cuda), p_b.to(cuda), s_w.to(cuda), s_b.to(cuda))
    got = ops.dynamic_infer(x.to(cuda), w_tap, b_cat, (3, 3), 1, n_valid=nv.to(cuda))
    for b in range(B):
        n = int(nv[b])
        ref, _ = O.din_ratio(x[b:b + 1, :, :n].contiguous(), p_w, p_b, s_w, s_b, (3, 3), 1)
        _close(got[b:b + 1, :, :n], ref, 2e-5, f"clip {b}")
        assert (got[b, :, n:] == 0).all()


def test_dpi_modules_match_oracle(cuda):
    import din_oracle as O
    from infer_module.dynamic_infer_module import (Dynamic_Person_Inference, Hierarchical_Dynamic_Inference,
                                                   Multi_Dynamic_Inference)
    g = torch.Generator().manual_seed(13)

    def randomise(mod):
        for n, p in mod.named_parameters():
            with torch.no_grad():
                if "p_conv" in n or "scale_conv" in n:
                    p.copy_(torch.randn(p.shape, generator=g) * (0.01 if n.endswith("weight") else 0.5))
                elif n.endswith("beta"):
                    p.copy_(torch.rand(p.shape, generator=g) + 0.5)
                elif "LN" in n:
                    p.copy_(torch.rand(p.shape, generator=g) + 0.5 if n.endswith("weight")
                            else torch.randn(p.shape, generator=g) * 0.1)
    # single module, two ratios, beta
    m = Dynamic_Person_Inference(128, (10, 12), kernel_size=(3, 3), dynamic_sampling=True, sampling_ratio=[1, 3],
                                 scale_factor=True, beta_factor=True)
    randomise(m)
    x = torch.randn(2, 10, 12, 128, generator=g)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    ref = O.dynamic_person_inference(x, sd, "", (3, 3), [1, 3], True, True)
    y, mad = m.to(cuda).eval()(x.to(cuda))
    _close(y, ref, 3e-5, "DPI")
    assert mad.numel() == 0
    # zero-initialised offsets/relation (the module's init state): DIN == zero-padded window mean
    m0 = Dynamic_Person_Inference(64, (10, 12), kernel_size=(3, 3), dynamic_sampling=True, sampling_ratio=[1],
                                  scale_factor=True)
    x0 = torch.randn(1, 4, 5, 64, generator=g)
    sd0 = {k: v.detach().clone() for k, v in m0.state_dict().items()}
    _close(m0.to(cuda).eval()(x0.to(cuda))[0], O.dynamic_person_inference(x0, sd0, "", (3, 3), [1], True, False),
           3e-5, "DPI zero-init")
    # parallel interaction fields (README "num_DIM=2")
    mm = Multi_Dynamic_Inference(64, (10, 12), kernel_size=[(1, 3), (3, 1)], dynamic_sampling=True,
                                 sampling_ratio=[1], scale_factor=True, num_DIM=2)
    randomise(mm)
    xm = torch.randn(2, 4, 5, 64, generator=g)
    sdm = {"DPI." + k: v.detach().clone() for k, v in mm.state_dict().items()}
    pc = O.PathConfig(ST_kernel_size=[(1, 3), (3, 1)], num_DIM=2, sampling_ratio=(1,))
    _close(mm.to(cuda).eval()(xm.to(cuda))[0], O.dpi_forward(xm, sdm, pc), 3e-5, "Multi DIM")
    # hierarchical (hard-coded 10 x 12 x 1024)
    mh = Hierarchical_Dynamic_Inference(1024, (10, 12), kernel_size=[(1, 3), (3, 1)], dynamic_sampling=True,
                                        sampling_ratio=[1], scale_factor=True)
    randomise(mh)
    xh = torch.randn(1, 10, 12, 1024, generator=g)
    sdh = {"DPI." + k: v.detach().clone() for k, v in mh.state_dict().items()}
    pch = O.PathConfig(ST_kernel_size=[(1, 3), (3, 1)], hierarchical_inference=True, sampling_ratio=(1,), lite_dim=None)
    _close(mh.to(cuda).eval()(xh.to(cuda))[0], O.dpi_forward(xh, sdh, pch), 5e-5, "Hierarchical")


def test_readout(cuda):
    from din_b200 import ops
    g = torch.Generator().manual_seed(14)
    B, T, N, C, A = 3, 4, 13, 128, 8
    s = torch.randn(B, T, N, C, generator=g)
    w, b = torch.randn(A, C, generator=g) * 0.1, torch.randn(A, generator=g)
    ref = F.linear(s.max(dim=2)[0], w, b).mean(dim=1)
    _close(ops.readout(s.to(cuda), w.to(cuda), b.to(cuda)), ref, 2e-5, "readout")
    nv = torch.tensor([13, 3, 1], dtype=torch.int32)
    ref = torch.stack([F.linear(s[i, :, :int(nv[i])].max(dim=1)[0], w, b).mean(dim=0) for i in range(B)])
    _close(ops.readout(s.to(cuda), w.to(cuda), b.to(cuda), n_valid=nv.to(cuda)), ref, 2e-5, "readout n_valid")


def test_dpi_modules_train_standalone(cuda):
    """Autograd through the stand-alone infer_module.dynamic_infer_module classes (reference
    dynamic_infer_module.py:121-151, 436-443, 491-498): gradient w.r.t. the input and every parameter against torch
    autograd over the oracle, for a bare module (two ratios, beta), parallel fields and the hierarchical pair."""
    import din_oracle as O
    from infer_module.dynamic_infer_module import (Dynamic_Person_Inference, Hierarchical_Dynamic_Inference,
                                                   Multi_Dynamic_Inference)
    g = torch.Generator().manual_seed(21)

    def randomise(mod):
        for n, p in mod.named_parameters():
            with torch.no_grad():
                if "p_conv" in n or "scale_conv" in n:
                    p.copy_(torch.randn(p.shape, generator=g) * (0.01 if n.endswith("weight") else 0.5))
                elif n.endswith("beta"):
                    p.copy_(torch.rand(p.shape, generator=g) + 0.5)
                elif "LN" in n:
                    p.copy_(torch.rand(p.shape, generator=g) + 0.5 if n.endswith("weight")
                            else torch.randn(p.shape, generator=g) * 0.1)

    def check(mod, x, ref_fn, prefix, what):
        randomise(mod)
        mod.eval()                                            # dropout off (hierarchical), gradients still flow
        sd = {prefix + k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in mod.state_dict().items()}
        xr = x.clone().requires_grad_(True)
        w = torch.randn(x.shape, generator=g)
        (ref_fn(xr, sd) * w).sum().backward()
        mod = mod.to(cuda)
        xc = x.to(cuda).requires_grad_(True)
        y, mad = mod(xc)
        assert y.requires_grad and mad.numel() == 0
        (y * w.to(cuda)).sum().backward()
        _close(xc.grad, xr.grad, 2e-4, what + " d/dx")
        n_checked = 0
        for n, p in mod.named_parameters():
            assert p.grad is not None, n
            _close(p.grad, sd[prefix + n].grad, 3e-4, f"{what} d/d{n}")
            n_checked += 1
        return n_checked

    m = Dynamic_Person_Inference(128, (10, 12), kernel_size=(3, 3), dynamic_sampling=True, sampling_ratio=[1, 3],
                                 scale_factor=True, beta_factor=True)
    n = check(m, torch.randn(2, 6, 7, 128, generator=g),
              lambda x, sd: O.dynamic_person_inference(x, sd, "", (3, 3), [1, 3], True, True), "", "DPI")
    assert n == 10                                            # hidden, beta, 2 x (p_conv w/b, scale_conv w/b)
    mm = Multi_Dynamic_Inference(64, (10, 12), kernel_size=[(1, 3), (3, 1)], dynamic_sampling=True,
                                 sampling_ratio=[1], scale_factor=True, num_DIM=2)
    pc = O.PathConfig(ST_kernel_size=[(1, 3), (3, 1)], num_DIM=2, sampling_ratio=(1,))
    check(mm, torch.randn(2, 4, 5, 64, generator=g), lambda x, sd: O.dpi_forward(x, sd, pc), "DPI.", "Multi")
    mh = Hierarchical_Dynamic_Inference(1024, (10, 12), kernel_size=[(1, 3), (3, 1)], dynamic_sampling=True,
                                        sampling_ratio=[1], scale_factor=True)
    pch = O.PathConfig(ST_kernel_size=[(1, 3), (3, 1)], hierarchical_inference=True, sampling_ratio=(1,), lite_dim=None)
    check(mh, torch.randn(1, 10, 12, 1024, generator=g), lambda x, sd: O.dpi_forward(x, sd, pch), "DPI.", "Hierarchical")
    # train() mode of the hierarchical module draws its dropout mask from torch's generator and still back-propagates
    mh.train()
    xt = torch.randn(1, 10, 12, 1024, generator=g).to(cuda).requires_grad_(True)
    yt, _ = mh(xt)
    yt.square().mean().backward()
    assert torch.isfinite(xt.grad).all() and float(xt.grad.abs().max()) > 0
