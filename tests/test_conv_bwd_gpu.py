"""Backward of the backbone convolutions (SURVEY.md §8f rank 1, VGG-16 slice) vs plain PyTorch fp32 autograd of
the same op on the SAME fp16-rounded operands (only the fp32 accumulation order differs)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

WGRAD_CASES = [
    # n, h, w, cin, cout, pad
    (1, 16, 8, 64, 64, 1),        # one tile, M half zero-filled (c_out = 64), NCI = 64
    (1, 16, 8, 128, 128, 1),      # one tile, NCI = 128
    (2, 45, 80, 128, 256, 1),     # ragged tiles (45 rows), two co blocks
    (1, 90, 160, 256, 256, 1),
    (3, 22, 40, 512, 512, 1),     # 4 x 4 x 3 units
    (1, 720, 1280, 64, 64, 1),    # VGG conv1_2 at full size: long K loop per CTA
    (1, 37, 29, 64, 192, 0),      # no padding, odd sizes, c_out = 1.5 blocks
]


@pytest.mark.parametrize("case", WGRAD_CASES, ids=[str(c) for c in WGRAD_CASES])
def test_wgrad_matches_torch(cuda, case):
    from din_b200 import ops
    n, h, w, cin, cout, pad = case
    g = torch.Generator().manual_seed(h * 7 + cin)
    x = torch.randn(n, h, w, cin, generator=g).to(cuda).half()
    oh, ow = h + 2 * pad - 2, w + 2 * pad - 2
    dz = (torch.randn(n, oh, ow, cout, generator=g) * 0.5).to(cuda).half()
    wt = torch.zeros(cout, cin, 3, 3, device=cuda, requires_grad=True)
    bias = torch.zeros(cout, device=cuda, requires_grad=True)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    y = F.conv2d(x.float().permute(0, 3, 1, 2), wt, bias, padding=pad)
    y.backward(dz.float().permute(0, 3, 1, 2))
    torch.backends.cudnn.allow_tf32 = old
    ref = wt.grad.permute(0, 2, 3, 1).contiguous()             # [co, kh, kw, ci]
    dw = torch.zeros(cout, 3, 3, cin, device=cuda)
    db = torch.zeros(cout, device=cuda)
    inv = torch.tensor([0.25], device=cuda)
    ops.conv2d_wgrad_nhwc(x, dz, dw, db, pad=(pad, pad), inv_scale=inv)
    torch.cuda.synchronize()
    err = (dw - 0.25 * ref).abs().max().item()
    scale = (0.25 * ref).abs().max().item()
    print(f"\n[wgrad] {case}: max|Δ| {err:.3e} max|ref| {scale:.3e}")
    assert err <= 1e-3 * scale, (err, scale)
    assert (db - 0.25 * bias.grad).abs().max().item() <= 1e-3 * (0.25 * bias.grad).abs().max().item()
    # accumulation: a second call adds on top
    ops.conv2d_wgrad_nhwc(x, dz, dw, db, pad=(pad, pad), inv_scale=inv)
    torch.cuda.synchronize()
    assert (dw - 0.5 * ref).abs().max().item() <= 1e-3 * (0.5 * ref).abs().max().item()


@pytest.mark.parametrize("shape", [(2, 16, 24, 64), (1, 45, 81, 128), (2, 7, 9, 8)], ids=str)
def test_relu_pool_bwd_matches_torch(cuda, shape):
    """fp16 activations quantise to few distinct values => plenty of exact ties inside 2x2 windows: the FIRST
    maximum in scan order must receive the gradient, as in torch."""
    from din_b200 import ops
    n, h, w, c = shape
    g = torch.Generator().manual_seed(h)
    pre = (torch.randn(n, h, w, c, generator=g) * 2).round() / 2               # multiples of 0.5: many ties
    y = F.relu(pre).to(cuda).half()
    for pool in (False, True):
        yy = y.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
        out = F.max_pool2d(yy, 2, 2) if pool else yy
        # relu's gradient mask from the saved OUTPUT (y > 0), as nn.ReLU(inplace=True) backward does
        dy = torch.randn(out.shape, generator=g).to(cuda).half()
        out.backward(dy.float())
        ref = (yy.grad * (yy.detach() > 0)).permute(0, 2, 3, 1)
        dz = ops.relu_pool_bwd_nhwc(y, dy.permute(0, 2, 3, 1).contiguous(), pool)
        torch.cuda.synchronize()
        assert torch.equal(dz.float(), ref), (shape, pool, (dz.float() - ref).abs().max().item())


def test_dgrad_is_forward_kernel_on_rotated_filter(cuda):
    """dX of a 3x3 s1 p1 conv = the forward tcgen05 kernel applied to dZ with the filter rotated by 180 degrees and
    its channel axes swapped; checked against torch's conv_transpose2d."""
    from din_b200 import ops
    g = torch.Generator().manual_seed(11)
    n, h, w, cin, cout = 2, 22, 40, 128, 256
    wt = (torch.randn(cout, cin, 3, 3, generator=g) * 0.05).to(cuda)
    dz = torch.randn(n, h, w, cout, generator=g).to(cuda).half()
    w_dgrad = ops.pack_conv_weight(wt.permute(1, 0, 2, 3).flip(2, 3).contiguous())
    dx = ops.conv2d_nhwc(dz, w_dgrad, None, stride=1, pad=(1, 1), relu=False)
    w_used = w_dgrad[..., :cout].float().permute(0, 3, 1, 2).flip(2, 3).permute(1, 0, 2, 3).contiguous()   # fp16-rounded OIHW
    ref = F.conv_transpose2d(dz.float().permute(0, 3, 1, 2), w_used, padding=1).permute(0, 2, 3, 1)
    torch.cuda.synchronize()
    assert (dx.float() - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()


@pytest.mark.parametrize("simt", [False, True], ids=["tensor", "simt"])
@pytest.mark.parametrize("u8", [False, True])
def test_stem_wgrad_matches_torch(cuda, u8, simt, monkeypatch):
    """Default = the tcgen05 kernel (fp16 im2col operands: 1e-3); DIN_STEM_WGRAD_SIMT=1 = the CUDA-core kernel kept
    for A/B measurements (fp32 patches)."""
    from din_b200 import ops
    monkeypatch.setenv("DIN_STEM_WGRAD_SIMT", "1" if simt else "0")
    g = torch.Generator().manual_seed(5)
    n, h, w = 2, 37, 150
    raw = torch.randint(0, 256, (n, 3, h, w), generator=g).float().to(cuda)
    dz = (torch.randn(n, h, w, 64, generator=g) * 0.5).to(cuda).half()
    wt = torch.zeros(64, 3, 3, 3, device=cuda, requires_grad=True)
    bias = torch.zeros(64, device=cuda, requires_grad=True)
    xp = ((raw / 255.0) - 0.5) * 2.0
    F.conv2d(xp, wt, bias, padding=1).backward(dz.float().permute(0, 3, 1, 2))
    dw = torch.zeros(64, 3, 3, 3, device=cuda)
    db = torch.zeros(64, device=cuda)
    inv = torch.tensor([0.5], device=cuda)
    x_in = raw.permute(0, 2, 3, 1).contiguous().to(torch.uint8) if u8 else raw
    ops.stem_wgrad(x_in, dz, dw, db, inv_scale=inv)
    torch.cuda.synchronize()
    # fp32 sums over 11 100 pixels in a different order than cuDNN's: 1.4e-4 measured
    assert (dw - 0.5 * wt.grad).abs().max().item() <= 1e-3 * (0.5 * wt.grad).abs().max().item()
    assert (db - 0.5 * bias.grad).abs().max().item() <= 1e-3 * (0.5 * bias.grad).abs().max().item()


def test_roi_align_bwd_matches_oracle_autograd(cuda):
    import din_oracle as O
    from din_b200 import ops
    g = torch.Generator().manual_seed(2)
    n_img, H, W, D, N = 3, 9, 14, 16, 5
    fm = torch.randn(n_img, D, H, W, generator=g, requires_grad=True)
    cx, cy = torch.rand(n_img * N, generator=g) * W, torch.rand(n_img * N, generator=g) * H
    bw, bh = 1 + 3 * torch.rand(n_img * N, generator=g), 2 + 5 * torch.rand(n_img * N, generator=g)
    boxes = torch.stack((cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2), dim=-1)     # some leave the map
    idx = torch.arange(n_img, dtype=torch.int32).repeat_interleave(N)
    out = O.roi_align_longcw(fm, boxes, idx, 5, 5)                                        # [M, D, 5, 5]
    dout = torch.randn(out.shape, generator=g)
    out.backward(dout)
    dcrops = dout.permute(0, 2, 3, 1).reshape(n_img * N, 25, D).contiguous().to(cuda)     # [m][bin][d]
    dfm = torch.zeros(n_img, H, W, D, device=cuda)
    ops.roi_align_bwd(dcrops, boxes.to(cuda), idx.to(cuda), dfm, 5, 5)
    torch.cuda.synchronize()
    ref = fm.grad.permute(0, 2, 3, 1)
    assert (dfm.cpu() - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()


def test_grad_to_f16_scale(cuda):
    from din_b200 import ops
    x = torch.randn(10000, device=cuda) * 3e-6
    y, ws = ops.grad_to_f16(x, target=256.0)
    torch.cuda.synchronize()
    s, inv = ws[1].item(), ws[2].item()
    amax = x.abs().max().item()
    assert s * inv == 1.0 and 128.0 <= amax * s <= 256.0 and (s == 2.0 ** round(torch.log2(torch.tensor(s)).item()))
    assert torch.equal(y, (x * s).half())
    z, ws = ops.grad_to_f16(torch.zeros(16, device=cuda))
    assert ws[1].item() == 1.0 and (z == 0).all()


# ------------------------------------------------------------------------------------------------
# ResNet-18 backward helpers
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(2, 12, 16, 64), (1, 45, 81, 64), (2, 9, 7, 8)], ids=str)
def test_maxpool3s2_relu_bwd_matches_torch(cuda, shape):
    from din_b200 import ops
    n, h, w, c = shape
    g = torch.Generator().manual_seed(w)
    x = F.relu((torch.randn(n, h, w, c, generator=g) * 2).round() / 2).to(cuda).half()        # ties + zeros
    xx = x.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    out = F.max_pool2d(xx, 3, 2, 1)
    dy = torch.randn(out.shape, generator=g).to(cuda).half()
    out.backward(dy.float())
    ref = (xx.grad * (xx.detach() > 0)).permute(0, 2, 3, 1)
    dz = ops.maxpool3s2_relu_bwd_nhwc(x, dy.permute(0, 2, 3, 1).contiguous())
    torch.cuda.synchronize()
    # a pixel can collect up to four windows' gradients: sums of fp16 values, rounded once more to fp16
    assert (dz.float() - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())
    assert ((dz.float() != 0) == (ref != 0)).float().mean().item() > 0.999


def test_scatter2_add_and_scale_rows(cuda):
    from din_b200 import ops
    g = torch.Generator().manual_seed(1)
    src = torch.randn(2, 5, 7, 16, generator=g).to(cuda).half()
    up = ops.scatter2_nhwc(src, 9, 13)
    ref = torch.zeros(2, 9, 13, 16, device=cuda, dtype=torch.float16)
    ref[:, ::2, ::2] = src
    assert torch.equal(up, ref)
    base = torch.randn(2, 10, 14, 16, generator=g).to(cuda).half()
    want = base.clone()
    want[:, :10:2, :14:2] += src
    ops.scatter2_nhwc(src, 10, 14, dst=base)
    assert torch.equal(base, want)
    a, b = torch.randn(3, 40, generator=g).to(cuda).half(), torch.randn(3, 40, generator=g).to(cuda).half()
    assert torch.equal(ops.add_f16(a, b), a + b)
    wgt = torch.randn(6, 3, 3, 8, generator=g).to(cuda)
    sc = torch.rand(6, generator=g).to(cuda) + 0.5
    want = wgt * sc.view(-1, 1, 1, 1)
    assert torch.allclose(ops.scale_rows(wgt.clone(), sc), want, rtol=1e-6, atol=0)


def test_stride2_conv_backward_by_zero_insertion(cuda):
    """dX and dW of a 3x3 stride-2 pad-1 convolution through the STRIDE-1 kernels on the zero-inserted dZ."""
    from din_b200 import ops
    g = torch.Generator().manual_seed(8)
    n, h, w, cin, cout = 2, 23, 40, 64, 128
    x = torch.randn(n, h, w, cin, generator=g).to(cuda).half()
    wt = (torch.randn(cout, cin, 3, 3, generator=g) * 0.05).to(cuda)
    oh, ow = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    dz = torch.randn(n, oh, ow, cout, generator=g).to(cuda).half()
    xr = x.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    w_dgrad = ops.pack_conv_weight(wt.permute(1, 0, 2, 3).flip(2, 3).contiguous())
    w_used = w_dgrad[..., :cout].float().permute(0, 3, 1, 2).flip(2, 3).permute(1, 0, 2, 3).contiguous().requires_grad_(True)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    F.conv2d(xr, w_used, None, stride=2, padding=1).backward(dz.float().permute(0, 3, 1, 2))
    torch.backends.cudnn.allow_tf32 = old
    dz_up = ops.scatter2_nhwc(dz, h, w)
    dx = ops.conv2d_nhwc(dz_up, w_dgrad, None, stride=1, pad=(1, 1), relu=False)
    dw = torch.zeros(cout, 3, 3, cin, device=cuda)
    ops.conv2d_wgrad_nhwc(x, dz_up, dw, None, pad=(1, 1))
    torch.cuda.synchronize()
    ref_dx = xr.grad.permute(0, 2, 3, 1)
    assert (dx.float() - ref_dx).abs().max().item() <= 2e-3 * ref_dx.abs().max().item()
    ref_dw = w_used.grad.permute(0, 2, 3, 1)
    assert (dw - ref_dw).abs().max().item() <= 1e-3 * ref_dw.abs().max().item()


def test_bn_gamma_grad_and_stem7_wgrad(cuda):
    from din_b200 import ops
    g = torch.Generator().manual_seed(12)
    # folded eval-mode BN + residual + ReLU: dgamma from the saved post-ReLU activation
    n, h, w, c = 2, 9, 11, 64
    conv = torch.randn(n, c, h, w, generator=g).to(cuda)
    idn = torch.randn(n, c, h, w, generator=g).to(cuda).half().float()
    gamma = (torch.rand(c, generator=g) + 0.5).to(cuda).requires_grad_(True)
    beta = (torch.randn(c, generator=g) * 0.2).to(cuda)
    mean, var = torch.randn(c, generator=g).to(cuda) * 0.1, torch.rand(c, generator=g).to(cuda) + 0.5
    z = F.batch_norm(conv, mean, var, gamma, beta, training=False, eps=1e-5)
    y = F.relu(z + idn)
    dy = torch.randn(y.shape, generator=g).to(cuda)
    y.backward(dy)
    yh = y.detach().permute(0, 2, 3, 1).contiguous().half()
    dzh = (dy * (y.detach() > 0)).permute(0, 2, 3, 1).contiguous().half()
    dgam = torch.zeros(c, device=cuda)
    ops.bn_gamma_grad(dzh, yh, gamma.detach(), beta, dgam, sub=idn.permute(0, 2, 3, 1).contiguous().half())
    torch.cuda.synchronize()
    assert (dgam - gamma.grad).abs().max().item() <= 5e-3 * gamma.grad.abs().max().item()
    # ResNet stem: 7x7 stride 2 pad 3 on prep(raw)
    raw = torch.randint(0, 256, (2, 3, 45, 150), generator=g).float().to(cuda)
    wt = torch.zeros(64, 3, 7, 7, device=cuda, requires_grad=True)
    bias = torch.zeros(64, device=cuda, requires_grad=True)
    out = F.conv2d(((raw / 255.0) - 0.5) * 2.0, wt, bias, stride=2, padding=3)
    dz = (torch.randn(out.shape, generator=g) * 0.5).to(cuda).half()
    out.backward(dz.float())
    for u8 in (False, True):
        dw, db = torch.zeros(64, 3, 7, 7, device=cuda), torch.zeros(64, device=cuda)
        xin = raw.permute(0, 2, 3, 1).contiguous().to(torch.uint8) if u8 else raw
        ops.stem_wgrad(xin, dz.permute(0, 2, 3, 1).contiguous(), dw, db, stride=2, pad=3)
        torch.cuda.synchronize()
        assert (dw - wt.grad).abs().max().item() <= 2e-3 * wt.grad.abs().max().item(), u8
        assert (db - bias.grad).abs().max().item() <= 2e-3 * bias.grad.abs().max().item(), u8


@pytest.mark.parametrize("ci,co,hw", [(64, 64, (40, 72)), (128, 128, (24, 40)), (512, 256, (12, 20)), (128, 64, (17, 23))],
                         ids=str)
def test_dgrad_with_fused_relu_backward_equals_conv_then_mask(cuda, ci, co, hw):
    """din_conv2d_relu_bwd_nhwc_f16 == din_conv2d_nhwc_f16 followed by the ReLU mask, bit for bit (every kernel variant:
    CTA pair BN = 64 / 128, one-CTA BN = 256)."""
    from din_b200 import ops
    g = torch.Generator().manual_seed(ci + co)
    h, w = hw
    dz = torch.randn(2, h, w, ci, generator=g).to(cuda).half()
    wp = ops.pack_conv_weight((torch.randn(co, ci, 3, 3, generator=g) * 0.05).to(cuda))
    y = torch.relu(torch.randn(2, h, w, co, generator=g)).to(cuda).half()
    plain = ops.conv2d_nhwc(dz, wp, None, stride=1, pad=(1, 1))
    fused = ops.conv2d_nhwc(dz, wp, None, stride=1, pad=(1, 1), relu_mask=y)
    torch.cuda.synchronize()
    want = torch.where(y > 0, plain, torch.zeros_like(plain))
    assert torch.equal(fused, want)
    assert (fused != 0).any() and (y == 0).any()


@pytest.mark.parametrize("shape", [(4, 45, 80, 64), (3, 12, 20, 512), (2, 7, 9, 24)], ids=str)
def test_batch_stat_batchnorm_forward_backward_match_torch(cuda, shape):
    """csrc/bn_train.cu vs F.batch_norm(training=True) + ReLU (+ residual) and its autograd, incl. running statistics."""
    from din_b200 import ops
    n, h, w, c = shape
    g = torch.Generator().manual_seed(c)
    z = (torch.randn(n, h, w, c, generator=g) * 1.5 + torch.randn(c, generator=g)).to(cuda).half()
    res = torch.randn(n, h, w, c, generator=g).to(cuda).half()
    gamma = (torch.rand(c, generator=g) + 0.5).to(cuda)
    beta = (torch.randn(c, generator=g) * 0.3).to(cuda)
    dy = torch.randn(n, h, w, c, generator=g).to(cuda).half()
    for residual, zin in ((None, z), (res, z), (res, z.float())):          # z as fp16 or fp32
        rm, rv = torch.zeros(c, device=cuda), torch.ones(c, device=cuda)
        rm_ref, rv_ref = rm.clone(), rv.clone()
        zz = z.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
        gg, bb = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
        pre = F.batch_norm(zz, rm_ref, rv_ref, gg, bb, training=True, momentum=0.1, eps=1e-5)
        if residual is not None:
            pre = pre + residual.float().permute(0, 3, 1, 2)
        ref = F.relu(pre)
        ref.backward(dy.float().permute(0, 3, 1, 2))
        y, (mean, invstd) = ops.bn_train_forward(zin, gamma, beta, rm, rv, residual=residual, relu=True)
        gmask = ops.relu_pool_bwd_nhwc(y, dy, False)
        dbeta, dgamma = torch.zeros(c, device=cuda), torch.zeros(c, device=cuda)
        dz = ops.bn_train_backward(gmask, zin, mean, invstd, gamma, dbeta, dgamma)
        torch.cuda.synchronize()
        assert (y.float() - ref.detach().permute(0, 2, 3, 1)).abs().max().item() <= 4e-3 * max(1.0, ref.abs().max().item())
        assert torch.allclose(rm, rm_ref, rtol=1e-4, atol=1e-5) and torch.allclose(rv, rv_ref, rtol=1e-4, atol=1e-5)

        def rel(a, b):
            return float((a.double() - b.double()).norm() / b.double().norm())
        # ReLU decisions of units within fp16 rounding of zero can differ from the fp32 reference
        assert rel(dz.float(), zz.grad.permute(0, 2, 3, 1)) <= 2e-2
        assert rel(dbeta, bb.grad) <= 1e-2 and rel(dgamma, gg.grad) <= 1e-2


# ---- Inception-v3 backward pieces (filter shapes / channel counts of backbone.py:10-85, concat slices) -----------------
WGRAD_GENERAL = [
    # n, h, w, cin, cout, (kh, kw), (ph, pw), x / dz channel offsets inside wider buffers
    (2, 17, 29, 192, 208, (1, 1), (0, 0), 0, 0),       # merged branch heads of Mixed_5b
    (1, 19, 23, 768, 704, (1, 1), (0, 0), 0, 0),       # merged heads of Mixed_6c: 6 c_out blocks x 6 NCI blocks
    (2, 17, 29, 48, 64, (5, 5), (2, 2), 64, 0),        # branch5x5_2: c_in = 48 (partial block), input = a slab slice
    (1, 19, 23, 160, 192, (1, 7), (0, 3), 0, 0),       # 1x7
    (1, 19, 23, 128, 128, (7, 1), (3, 0), 128, 0),     # 7x1 reading a slab slice
    (1, 30, 40, 80, 192, (3, 3), (0, 0), 0, 0),        # Conv2d_4a: c_in = 80, no padding
    (1, 37, 41, 32, 32, (3, 3), (0, 0), 0, 0),         # Conv2d_2a: 32 -> 32
    (1, 21, 27, 96, 96, (3, 3), (1, 1), 0, 32),        # dz = a slice of a wider gradient buffer
    (1, 16, 24, 288, 64, (1, 1), (0, 0), 0, 0),        # Mixed_6a branch3x3dbl_1: c_in = 288 (4.5 blocks)
]


@pytest.mark.parametrize("case", WGRAD_GENERAL, ids=[str(c[3:7]) for c in WGRAD_GENERAL])
def test_wgrad_general_filters_and_channel_slices(cuda, case):
    from din_b200 import ops
    n, h, w, cin, cout, (kh, kw), (ph, pw), xo, dzo = case
    g = torch.Generator().manual_seed(h * 7 + cin + kw)
    xb = torch.randn(n, h, w, cin + xo + 8, generator=g).to(cuda).half()
    oh, ow = h + 2 * ph - kh + 1, w + 2 * pw - kw + 1
    dzb = (torch.randn(n, oh, ow, cout + dzo + 16, generator=g) * 0.5).to(cuda).half()
    x, dz = xb[..., xo:xo + cin], dzb[..., dzo:dzo + cout]
    wt = torch.zeros(cout, cin, kh, kw, device=cuda, requires_grad=True)
    bias = torch.zeros(cout, device=cuda, requires_grad=True)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    F.conv2d(x.float().permute(0, 3, 1, 2), wt, bias, padding=(ph, pw)).backward(dz.float().permute(0, 3, 1, 2))
    torch.backends.cudnn.allow_tf32 = old
    ref = wt.grad.permute(0, 2, 3, 1).contiguous()
    dw = torch.zeros(cout, kh, kw, cin, device=cuda)
    db = torch.zeros(cout, device=cuda)
    ops.conv2d_wgrad_nhwc(xb, dzb, dw, db, pad=(ph, pw), x_c_offset=xo, dz_c_offset=dzo)
    torch.cuda.synchronize()
    err, scale = (dw - ref).abs().max().item(), ref.abs().max().item()
    print(f"\n[wgrad general] {case}: max|Δ| {err:.3e} max|ref| {scale:.3e}")
    assert err <= 1e-3 * scale, (err, scale)
    assert (db - bias.grad).abs().max().item() <= 1e-3 * bias.grad.abs().max().item()


@pytest.mark.parametrize("u8", [False, True])
def test_stem_wgrad_inception_geometry(cuda, u8):
    """Conv2d_1a_3x3: 32 output channels, stride 2, no padding (the dZ box is half out of bounds: zero-filled rows of M)."""
    from din_b200 import ops
    g = torch.Generator().manual_seed(6)
    n, h, w = 2, 37, 151
    raw = torch.randint(0, 256, (n, 3, h, w), generator=g).float().to(cuda)
    oh, ow = (h - 3) // 2 + 1, (w - 3) // 2 + 1
    dz = (torch.randn(n, oh, ow, 32, generator=g) * 0.5).to(cuda).half()
    wt = torch.zeros(32, 3, 3, 3, device=cuda, requires_grad=True)
    bias = torch.zeros(32, device=cuda, requires_grad=True)
    F.conv2d(((raw / 255.0) - 0.5) * 2.0, wt, bias, stride=2).backward(dz.float().permute(0, 3, 1, 2))
    dw, db = torch.zeros(32, 3, 3, 3, device=cuda), torch.zeros(32, device=cuda)
    x_in = raw.permute(0, 2, 3, 1).contiguous().to(torch.uint8) if u8 else raw
    ops.stem_wgrad(x_in, dz, dw, db, stride=2, pad=0)
    torch.cuda.synchronize()
    assert (dw - wt.grad).abs().max().item() <= 1e-3 * wt.grad.abs().max().item()
    assert (db - bias.grad).abs().max().item() <= 1e-3 * bias.grad.abs().max().item()


@pytest.mark.parametrize("shape", [(2, 13, 17, 64), (1, 44, 81, 192), (2, 9, 8, 288)], ids=str)
def test_maxpool3s2_pad0_bwd_slices_and_accumulate(cuda, shape):
    """F.max_pool2d(x, 3, 2) backward over channel slices of wider buffers, accumulating into an existing gradient."""
    from din_b200 import ops
    n, h, w, c = shape
    g = torch.Generator().manual_seed(w)
    xb = F.relu((torch.randn(n, h, w, c + 32, generator=g) * 2).round() / 2).to(cuda).half()      # ties + zeros
    x = xb[..., 8:8 + c]
    xx = x.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    out = F.max_pool2d(xx, 3, 2)
    dyb = torch.randn(n, out.shape[2], out.shape[3], c + 16, generator=g).to(cuda).half()
    out.backward(dyb[..., 16:16 + c].float().permute(0, 3, 1, 2))
    ref = (xx.grad * (xx.detach() > 0)).permute(0, 2, 3, 1)
    prev = torch.randn(n, h, w, c + 8, generator=g).to(cuda).half()
    dz = prev.clone()
    ops.maxpool3s2_bwd_nhwc(xb, dyb, dz, c=c, pad=0, x_c_offset=8, dy_c_offset=16, dz_c_offset=0, accumulate=True)
    torch.cuda.synchronize()
    want = prev[..., :c].float() + ref
    assert (dz[..., :c].float() - want).abs().max().item() <= 4e-3 * max(1.0, want.abs().max().item())
    assert torch.equal(dz[..., c:], prev[..., c:])
    dz2 = torch.zeros(n, h, w, c, dtype=torch.float16, device=cuda)
    ops.maxpool3s2_bwd_nhwc(xb, dyb, dz2, c=c, pad=0, x_c_offset=8, dy_c_offset=16)
    torch.cuda.synchronize()
    assert (dz2.float() - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())
    assert ((dz2.float() != 0) == (ref != 0)).float().mean().item() > 0.999


def test_relu_slice_upsample_bwd_and_colsum(cuda):
    from din_b200 import ops
    g = torch.Generator().manual_seed(4)
    n, h, w, c = 2, 9, 14, 48
    y = F.relu(torch.randn(n, h, w, 96, generator=g)).to(cuda).half()
    dy = torch.randn(n, h, w, 64, generator=g).to(cuda).half()
    dz = torch.full((n, h, w, 80), 3.0, dtype=torch.float16, device=cuda)
    ops.relu_bwd_slice_nhwc(y, dy, dz, c=c, y_c_offset=32, dy_c_offset=8, dz_c_offset=16)
    torch.cuda.synchronize()
    assert torch.equal(dz[..., 16:16 + c], dy[..., 8:8 + c] * (y[..., 32:32 + c] > 0))
    assert (dz[..., :16] == 3).all() and (dz[..., 16 + c:] == 3).all()
    # column sums of a slice
    db = torch.zeros(c, device=cuda)
    ops.colsum_nhwc(dy, db, c=c, c_offset=8, inv_scale=torch.tensor([0.5], device=cuda))
    torch.cuda.synchronize()
    want = 0.5 * dy[..., 8:8 + c].float().sum(dim=(0, 1, 2))
    assert (db - want).abs().max().item() <= 1e-4 * max(1.0, want.abs().max().item())
    # adjoint of the align_corners=True bilinear resize (Mixed_6e -> the multiscale map), incl. a non-2x ratio
    for (hh, ww, oh, ow) in ((7, 11, 15, 23), (5, 6, 13, 17), (4, 4, 4, 4)):
        src = torch.randn(n, 64, hh, ww, generator=g).to(cuda).requires_grad_(True)
        up = F.interpolate(src, size=(oh, ow), mode="bilinear", align_corners=True)
        gb = torch.randn(n, oh, ow, 64 + 24, generator=g).to(cuda).half()
        up.backward(gb[..., 24:].float().permute(0, 3, 1, 2))
        dx = ops.upsample_bilinear_bwd_nhwc(gb, hh, ww, c=64, dy_c_offset=24)
        torch.cuda.synchronize()
        ref = src.grad.permute(0, 2, 3, 1)
        assert (dx.float() - ref).abs().max().item() <= 2e-3 * ref.abs().max().item(), (hh, ww, oh, ow)


def test_general_dgrad_filters(cuda):
    """Data gradients of the Inception filter shapes = the forward kernel on the transposed, rotated filter with padding
    k - 1 - pad; stride-2 pad-0 layers through zero insertion (Mixed_6a)."""
    from din_b200 import ops
    g = torch.Generator().manual_seed(9)
    for (cin, cout, k, pad, stride, hw) in ((48, 64, (5, 5), (2, 2), 1, (17, 21)), (160, 192, (1, 7), (0, 3), 1, (12, 19)),
                                            (128, 128, (7, 1), (3, 0), 1, (12, 19)), (80, 192, (3, 3), (0, 0), 1, (15, 22)),
                                            (96, 96, (3, 3), (0, 0), 2, (17, 23)), (208, 192, (1, 1), (0, 0), 1, (9, 13))):
        kh, kw = k
        h, w = hw
        wt = (torch.randn(cout, cin, kh, kw, generator=g) * (2.0 / (cin * kh * kw)) ** 0.5).to(cuda)
        wd = ops.pack_conv_weights([(wt, None, 1, True)])[0]
        x = torch.randn(2, cin, h, w, generator=g).to(cuda).requires_grad_(True)
        y = F.conv2d(x, wd[..., :cout].float().permute(3, 0, 1, 2).flip(2, 3).contiguous(), stride=stride, padding=pad)
        dz = torch.randn(y.shape, generator=torch.Generator().manual_seed(1)).to(cuda).half()
        y.backward(dz.float())
        dzn = dz.permute(0, 2, 3, 1).contiguous()
        if stride == 2:
            dzn = ops.scatter2_nhwc(dzn, h - kh + 1 + 2 * pad[0], w - kw + 1 + 2 * pad[1])
        dx = ops.conv2d_nhwc(dzn, wd, None, stride=1, pad=(kh - 1 - pad[0], kw - 1 - pad[1]), relu=False, c_in=cout)
        torch.cuda.synchronize()
        ref = x.grad.permute(0, 2, 3, 1)
        assert dx.shape == ref.shape
        assert (dx.float() - ref).abs().max().item() <= 2e-3 * ref.abs().max().item(), (cin, cout, k)


@pytest.mark.parametrize("shape", [(64, 32, 3, 3), (128, 64, 1, 1), (64, 3, 7, 7), (192, 160, 1, 7)], ids=str)
def test_bn_fold_grads_matches_autograd(cuda, shape):
    """din_bn_fold_grads_f32 + din_scale_rows_f32 in isolation: conv -> eval-mode BatchNorm, gradients of the conv weight
    and of gamma from the gradient of the FOLDED weight, vs fp32 autograd (non-trivial running statistics, small gammas)."""
    from din_b200 import ops
    co, ci, kh, kw = shape
    g = torch.Generator().manual_seed(co + kw)
    x = torch.randn(2, ci, 9, 11, generator=g).to(cuda)
    w = (torch.randn(co, ci, kh, kw, generator=g) * 0.1).to(cuda).requires_grad_(True)
    gamma = (torch.rand(co, generator=g) * 1.5 + 1e-3).to(cuda).requires_grad_(True)     # down to ~1e-3: no division by gamma
    beta = torch.randn(co, generator=g).to(cuda).requires_grad_(True)
    mean, var = torch.randn(co, generator=g).to(cuda), (torch.rand(co, generator=g) + 0.5).to(cuda)
    eps = 1e-3
    old_tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    y = F.batch_norm(F.conv2d(x, w, padding=(kh // 2, kw // 2)), mean, var, gamma, beta, False, 0.0, eps)
    dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(2)).to(cuda)
    y.backward(dy)
    # what the CUDA path has: gradients of the folded convolution z = conv(x, scale * w) + shift
    scale = (gamma / torch.sqrt(var + eps)).detach()
    wf = (w.detach() * scale.view(-1, 1, 1, 1)).requires_grad_(True)
    shift = (beta.detach() - mean * scale).requires_grad_(True)
    (F.conv2d(x, wf, padding=(kh // 2, kw // 2)) + shift.view(1, -1, 1, 1)).backward(dy)
    torch.backends.cudnn.allow_tf32 = old_tf32
    dwf, dbeta = wf.grad.contiguous(), shift.grad.contiguous()
    dgamma = torch.zeros(co, device=cuda)
    ops.bn_fold_grads(w.detach().contiguous(), dwf, dbeta, mean, var, dgamma, eps=eps)
    dw = ops.scale_rows(dwf.clone(), scale.contiguous())
    torch.cuda.synchronize()
    assert (dgamma - gamma.grad).abs().max().item() <= 2e-5 * gamma.grad.abs().max().item()
    assert (dw - w.grad).abs().max().item() <= 2e-5 * w.grad.abs().max().item()
    assert (dbeta - beta.grad).abs().max().item() <= 2e-5 * beta.grad.abs().max().item()
