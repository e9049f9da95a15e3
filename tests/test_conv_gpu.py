"""tcgen05 implicit-GEMM convolution vs a plain PyTorch fp32 reference of the same op (GPU).

Both sides see the SAME fp16-rounded operands, so the only difference is fp32 accumulation order:
tolerance 2e-3 * max|ref| on the fp16 output (one fp16 ulp of the largest value is 9.8e-4).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _unpack(wp, cin):
    """packed fp16 [co][kh][kw][cin_p] -> OIHW fp32: the exact weights the kernel multiplies with."""
    return wp[..., :cin].float().permute(0, 3, 1, 2).contiguous()


def _ref_conv(x_nhwc_h, w_oihw_h, bias, stride, pad, relu, residual=None):
    x = x_nhwc_h.float().permute(0, 3, 1, 2)
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        y = F.conv2d(x, w_oihw_h.float(), bias, stride=stride, padding=pad)
        torch.backends.cuda.matmul.allow_tf32 = old
    if residual is not None:
        y = y + residual.float().permute(0, 3, 1, 2)
    if relu:
        y = F.relu(y)
    return y.permute(0, 2, 3, 1).contiguous()


CASES = [
    # n, h, w, cin, cout, k, stride, pad, relu
    (1, 8, 16, 64, 64, 1, 1, 0, False),      # one tile, one K step, single tap
    (1, 8, 16, 64, 64, 3, 1, 1, False),      # 3x3, halo zero-fill on all sides
    (2, 45, 80, 128, 128, 3, 1, 1, True),    # ragged tiles (45 rows), BN=128
    (1, 90, 160, 256, 256, 3, 1, 1, True),   # BN=256
    (1, 22, 40, 512, 512, 3, 1, 1, True),    # two N tiles, deep K (72 steps)
    (3, 36, 64, 64, 128, 3, 2, 1, True),     # stride 2 (ResNet downsampling 3x3)
    (2, 36, 64, 64, 128, 1, 2, 0, False),    # stride-2 1x1 (ResNet shortcut)
    (1, 23, 40, 192, 320, 1, 1, 0, True),    # cout not a multiple of BN, cin = 3 blocks
    (1, 17, 29, 128, 192, (1, 7), 1, (0, 3), True),   # Inception 1x7
    (1, 17, 29, 128, 192, (7, 1), 1, (3, 0), True),   # Inception 7x1
    (1, 720, 1280, 64, 64, 3, 1, 1, True),   # VGG conv1_2 at full size (7200 tiles, persistent loop)
    (1, 35, 35, 48, 64, 5, 1, 2, True),      # Inception 5x5, c_in = 48: partial K block zero-filled by TMA
    (1, 30, 40, 80, 192, 3, 1, 0, True),     # Inception Conv2d_4a: c_in = 80, no padding, BN = 192 tile
    (2, 17, 17, 160, 96, (7, 1), 1, (3, 0), True),   # c_in = 160, BN = 96 tile
    (1, 35, 35, 288, 384, 3, 2, 0, True),    # Mixed_6a: stride 2, pad 0, 288 -> 384 (two BN = 192 tiles)
    (1, 37, 37, 32, 32, 3, 1, 0, True),      # Conv2d_2a: 32 -> 32
]


@pytest.mark.parametrize("case", CASES, ids=[str(c) for c in CASES])
def test_conv_matches_torch(cuda, case):
    from din_b200 import ops
    n, h, w, cin, cout, k, stride, pad, relu = case
    kh, kw = (k, k) if isinstance(k, int) else k
    ph, pw = (pad, pad) if isinstance(pad, int) else pad
    g = torch.Generator(device="cpu").manual_seed(1234)
    x = (torch.randn(n, h, w, cin, generator=g) * 1.0).to(cuda).half()
    wt = (torch.randn(cout, cin, kh, kw, generator=g) * (2.0 / (cin * kh * kw)) ** 0.5).to(cuda)
    bias = torch.randn(cout, generator=g).to(cuda)
    wp = ops.pack_conv_weight(wt)
    assert wp.shape == (cout, kh, kw, (cin + 63) // 64 * 64)
    # packed weight: tap-major, zero K padding, error-feedback rounding: every element within one fp16 ulp
    # of the row's largest weight (own rounding + the carried residual) and each output row's SUMMED rounding
    # error ~1 ulp instead of ~sqrt(K) ulps
    wu = _unpack(wp, cin)
    row_ulp = wt.abs().amax(dim=(1, 2, 3), keepdim=True) * 2.0 ** -10
    assert ((wu - wt).abs() <= row_ulp).all()
    assert ((wu - wt).sum(dim=(1, 2, 3)).abs() <= 2 * row_ulp.flatten()).all()
    assert (wp[..., cin:] == 0).all()
    y = ops.conv2d_nhwc(x, wp, bias, stride=stride, pad=(ph, pw), relu=relu)
    torch.cuda.synchronize()
    ref = _ref_conv(x, wu, bias, stride, (ph, pw), relu)
    assert y.shape == ref.shape
    err = (y.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-3 * scale, f"max abs err {err} vs scale {scale}"


def test_conv_f32_out_residual_and_channel_slices(cuda):
    from din_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(7)
    n, h, w = 2, 23, 40
    xbuf = torch.randn(n, h, w, 192, generator=g).to(cuda).half()      # conv reads channels [64,192)
    wt = (torch.randn(72, 128, 3, 3, generator=g) * 0.03).to(cuda)
    bias = torch.randn(72, generator=g).to(cuda)
    res = torch.randn(n, h, w, 200, generator=g).to(cuda).half()
    out = torch.zeros(n, h, w, 200, device=cuda, dtype=torch.float16)  # conv writes channels [64,136)
    wp = ops.pack_conv_weight(wt)
    ops.conv2d_nhwc(xbuf, wp, bias, stride=1, pad=(1, 1), relu=True, residual=res, out=out,
                    c_in=128, x_c_offset=64, y_c_offset=64)
    ref = _ref_conv(xbuf[..., 64:192].contiguous(), _unpack(wp, 128), bias, 1, (1, 1), True,
                    residual=res[..., 64:136].contiguous())
    torch.cuda.synchronize()
    assert (out[..., :64] == 0).all() and (out[..., 136:] == 0).all()   # neighbours untouched
    err = (out[..., 64:136].float() - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item()
    # fp32 output (the fc_emb GEMM shape: rows x K -> 1024), tighter tolerance
    a = torch.randn(1, 1, 360, 1280, generator=g).to(cuda).half()
    wl = (torch.randn(1024, 1280, 1, 1, generator=g) * 0.03).to(cuda)
    bl = torch.randn(1024, generator=g).to(cuda)
    wlp = ops.pack_conv_weight(wl)
    y = ops.conv2d_nhwc(a, wlp, bl, out_f32=True)
    ref = a.float().reshape(360, 1280) @ _unpack(wlp, 1280).reshape(1024, 1280).t() + bl
    torch.cuda.synchronize()
    err = (y.reshape(360, 1024) - ref).abs().max().item()
    assert err <= 2e-5 * max(1.0, ref.abs().max().item()) * 10, err


@pytest.mark.parametrize("case", [(2, 45, 80, 128, 128, 3, 1, (1, 1)), (1, 23, 40, 192, 320, 1, 1, (0, 0)),
                                  (1, 35, 35, 288, 384, 3, 2, (0, 0)), (1, 17, 29, 160, 192, (1, 7), 1, (0, 3))],
                         ids=str)
def test_conv_split_weights(cuda, case):
    """w_split = 2: hi + lo fp16 weight parts accumulated in one TMEM tile == convolution with the (nearly)
    exact fp32 weights; tolerance 5e-4 (the fp16 rounding of the output)."""
    from din_b200 import ops
    n, h, w, cin, cout, k, stride, pad = case
    kh, kw = (k, k) if isinstance(k, int) else k
    g = torch.Generator(device="cpu").manual_seed(99)
    x = torch.randn(n, h, w, cin, generator=g).to(cuda).half()
    wt = (torch.randn(cout, cin, kh, kw, generator=g) * (2.0 / (cin * kh * kw)) ** 0.5).to(cuda)
    bias = torch.randn(cout, generator=g).to(cuda)
    wp = ops.pack_conv_weight(wt, split=2)
    assert wp.shape == (cout, 2, kh, kw, (cin + 63) // 64 * 64)
    rec = (wp[:, 0, ..., :cin].float() + wp[:, 1, ..., :cin].float()).permute(0, 3, 1, 2)
    assert (rec - wt).abs().max().item() <= 2.0 ** -20 * wt.abs().max().item()
    y = ops.conv2d_nhwc(x, wp, bias, stride=stride, pad=pad, relu=True, out_f32=True)
    ref = _ref_conv(x, wt, bias, stride, pad, True)
    torch.cuda.synchronize()
    assert (y - ref).abs().max().item() <= 2e-5 * ref.abs().max().item() * 4


def test_stem_and_pool(cuda):
    from din_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(3)
    x = torch.randint(0, 256, (2, 3, 70, 150), generator=g).float().to(cuda)
    for (co, k, s, p) in [(64, 3, 1, 1), (64, 7, 2, 3), (32, 3, 2, 0)]:
        wt = (torch.randn(co, 3, k, k, generator=g) * 0.1).to(cuda)
        b = torch.randn(co, generator=g).to(cuda)
        y = ops.stem_conv(x, wt, b, stride=s, pad=p, relu=True, prep=True)
        xp = ((x / 255.0) - 0.5) * 2.0
        ref = F.relu(F.conv2d(xp, wt, b, stride=s, padding=p)).permute(0, 2, 3, 1)
        torch.cuda.synchronize()
        assert y.shape == ref.shape
        err = (y.float() - ref).abs().max().item()
        assert err <= 1e-3 * ref.abs().max().item() + 1e-3, (co, k, s, p, err)
    t = torch.randn(2, 45, 81, 64, generator=g).to(cuda).half()
    for (k, s, p) in [(2, 2, 0), (3, 2, 1), (3, 2, 0)]:
        y = ops.maxpool2d_nhwc(t, k, s, p)
        ref = F.max_pool2d(t.float().permute(0, 3, 1, 2), k, s, p).permute(0, 2, 3, 1)
        torch.cuda.synchronize()
        assert torch.equal(y.float(), ref), (k, s, p)
    # channel-sliced pools writing into a concat buffer (Inception Mixed_6a), average pool, bilinear upsample
    out = torch.zeros(2, 22, 40, 96, device=cuda, dtype=torch.float16)
    ops.maxpool2d_nhwc(t, 3, 2, 0, out=out, c=48, x_c_offset=8, y_c_offset=40)
    ref = F.max_pool2d(t[..., 8:56].float().permute(0, 3, 1, 2), 3, 2).permute(0, 2, 3, 1)
    assert torch.equal(out[..., 40:88].float(), ref) and (out[..., :40] == 0).all() and (out[..., 88:] == 0).all()
    y = ops.avgpool2d_nhwc(t, 3, 1, 1)
    ref = F.avg_pool2d(t.float().permute(0, 3, 1, 2), 3, 1, 1).permute(0, 2, 3, 1)
    assert (y.float() - ref).abs().max().item() <= 2e-3
    up = torch.zeros(2, 91, 163, 80, device=cuda, dtype=torch.float16)
    ops.upsample_bilinear_nhwc(t, 91, 163, out=up, c=64, y_c_offset=16)
    ref = F.interpolate(t.float().permute(0, 3, 1, 2), size=(91, 163), mode="bilinear", align_corners=True)
    assert (up[..., 16:].float() - ref.permute(0, 2, 3, 1)).abs().max().item() <= 4e-3
    assert (up[..., :16] == 0).all()


@pytest.mark.parametrize("shape", [(64, 3, 3, 3), (128, 64, 3, 3), (512, 512, 3, 3), (1024, 12800, 1, 1), (40, 72, 1, 7),
                                   (64, 3, 7, 7)], ids=str)
def test_weight_packing_fast_kernel_is_bit_identical_to_the_serial_one(cuda, shape, monkeypatch):
    """The shared-memory-staged packing kernel (re-run on every training step) must reproduce the element-by-element
    error-feedback kernel bit for bit, with and without a BN scale, for both weight formats."""
    from din_b200 import ops
    g = torch.Generator().manual_seed(sum(shape))
    w = (torch.randn(*shape, generator=g) * 0.05).to(cuda)
    sc = (torch.rand(shape[0], generator=g) + 0.5).to(cuda)
    for scale in (None, sc):
        for split in (1, 2):
            monkeypatch.setenv("DIN_PACK_SERIAL", "1")
            ref = ops.pack_conv_weight(w, scale, split=split)
            monkeypatch.setenv("DIN_PACK_SERIAL", "0")
            out = ops.pack_conv_weight(w, scale, split=split)
            torch.cuda.synchronize()
            assert out.shape == ref.shape and torch.equal(out, ref), (shape, scale is not None, split)


def test_batched_packing_matches_single_packs_and_dgrad_filters(cuda):
    """din_pack_conv_weights_f16: one launch for many weights == one din_pack_conv_weight_f16 each (bit for bit), and its
    transposed mode == packing w.permute(1,0,2,3).flip(2,3) (x the BN scale), the data-gradient filter."""
    from din_b200 import ops
    g = torch.Generator().manual_seed(5)
    shapes = [(64, 64, 3, 3), (128, 64, 3, 3), (256, 128, 3, 3), (128, 64, 1, 1), (512, 512, 3, 3), (24, 40, 1, 7),
              (1024, 3200, 1, 1)] + [(64, 64, 3, 3)] * 70              # > 64 jobs: more than one launch
    ws = [(torch.randn(*s, generator=g) * 0.05).to(cuda) for s in shapes]
    scs = [(torch.rand(s[0], generator=g) + 0.5).to(cuda) if i % 2 else None for i, s in enumerate(shapes)]
    splits = [2 if i == 3 else 1 for i in range(len(shapes))]
    outs = ops.pack_conv_weights([(w, sc, sp, False) for w, sc, sp in zip(ws, scs, splits)])
    for w, sc, sp, out in zip(ws, scs, splits, outs):
        assert torch.equal(out, ops.pack_conv_weight(w, sc, split=sp)), tuple(w.shape)
    outs = ops.pack_conv_weights([(w, sc, 1, True) for w, sc in zip(ws[:7], scs[:7])])
    for w, sc, out in zip(ws[:7], scs[:7], outs):
        wf = w if sc is None else w * sc.view(-1, 1, 1, 1)
        ref = ops.pack_conv_weight(wf.permute(1, 0, 2, 3).flip(2, 3).contiguous())
        assert out.shape == ref.shape and torch.equal(out, ref), tuple(w.shape)


@pytest.mark.parametrize("n,h,w,pool,u8", [(2, 48, 64, True, False), (1, 50, 76, False, False), (3, 37, 100, True, False),
                                            (2, 48, 64, True, True), (1, 66, 112, False, True),
                                            (2, 720, 1280, True, False), (1, 720, 1280, True, True)],
                         ids=lambda v: str(v))
def test_stem_pair_fused_equals_the_two_kernel_path(cuda, n, h, w, pool, u8):
    """conv1_1 + conv1_2 in one launch (conv1_fused_2cta_kernel: the stem computed inside conv1_2's A producer) must
    reproduce the stand-alone stem followed by the CTA-pair convolution: conv1_1's values are bit-identical (same
    im2col, same K order, same fp16 rounding), conv1_2's differ only in WHERE its bias joins the fp32 sum (first, on the
    tensor core, instead of last, in the epilogue), i.e. by at most one fp16 ulp on a small fraction of the elements
    (with the bias in the epilogue the two paths were bit-identical on all 45 M elements of these cases) -- and both
    match fp32 torch.  Shapes cover
    ragged tiles (h % 16, w % 8 != 0), odd tile counts (an idle second CTA in the last pair), uint8 frames, 720p."""
    from din_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(11)
    img = torch.randint(0, 256, (n, 3, h, w), generator=g)
    w1 = (torch.randn(64, 3, 3, 3, generator=g) * 0.3).to(cuda)
    b1 = (torch.randn(64, generator=g) * 0.2).to(cuda)
    w2 = (torch.randn(64, 64, 3, 3, generator=g) * (2.0 / 576) ** 0.5).to(cuda)
    b2 = (torch.randn(64, generator=g) * 0.1).to(cuda)
    w2p = ops.pack_conv_weight(w2)
    x = img.permute(0, 2, 3, 1).contiguous().to(torch.uint8).to(cuda) if u8 else img.float().to(cuda)
    assert ops.stem_pair_supported(x)
    y1 = ops.stem_conv(x, w1, b1, stride=1, pad=1, relu=True, prep=True)
    want = ops.conv2d_nhwc(y1, w2p, b2, stride=1, pad=(1, 1), relu=True, pool2=pool)
    got = ops.stem_conv_pair(x, w1, b1, w2p, b2, relu=True, pool2=pool, prep=True)
    torch.cuda.synchronize()
    assert got.shape == want.shape
    diff = (got.float() - want.float()).abs()
    n_diff = int((diff > 0).sum())
    print(f"\n[stem pair] {n}x{h}x{w} pool={pool} u8={u8}: {n_diff} of {diff.numel()} elements differ, max |diff| {diff.max().item():.3e}")
    # one fp16 ulp at each element's magnitude, or -- for outputs near zero, where an fp16 ulp is far below the fp32
    # rounding of a 577-term sum of O(1) terms -- 1e-5 absolute
    tol = (want.float().abs() * 2.0 ** -10).clamp_min(1e-5)
    assert n_diff <= 2e-3 * diff.numel(), n_diff
    assert bool((diff <= tol).all()), float((diff / tol).max())
    if h * w <= 128 * 128:
        xp = ((img.float().to(cuda) / 255.0) - 0.5) * 2.0
        ref = F.relu(F.conv2d(F.relu(F.conv2d(xp, w1, b1, padding=1)), w2, b2, padding=1))
        if pool:
            ref = F.max_pool2d(ref, 2, 2)
        ref = ref.permute(0, 2, 3, 1)
        assert (got.float() - ref).abs().max().item() <= 4e-3 * ref.abs().max().item()


def test_stem_pair_rejects_unaligned_widths(cuda):
    from din_b200 import _lib, ops
    x = torch.zeros(1, 3, 32, 30, device=cuda)
    assert not ops.stem_pair_supported(x)
    w2p = ops.pack_conv_weight(torch.zeros(64, 64, 3, 3, device=cuda))
    with pytest.raises(_lib.DinError):
        ops.stem_conv_pair(x, torch.zeros(64, 3, 3, 3, device=cuda), None, w2p, None)


@pytest.mark.parametrize("shape", [(2, 21, 37, 192, (64, 64, 32, 48), 64, (128, 160)),      # Mixed_5b heads
                                   (1, 11, 19, 768, (192, 160, 160, 192), 192, (512, 704)),   # Mixed_6c heads: 3 N tiles
                                   (1, 5, 7, 288, (64, 64, 64, 48), 64, (128, 192))],
                         ids=["5b", "6c", "5d"])
def test_branch_group_gemm_and_pool_tail(cuda, shape):
    """din_conv2d_branches_nhwc_f16: stacked 1x1 convolutions in one launch, two destinations, a column range without
    ReLU; din_avgpool3_bias_relu_nhwc_f16 on that range equals conv1x1(avg_pool(x)) + bias, ReLU (the reference's
    branch_pool order, torchvision InceptionA/C.forward) up to fp16 rounding."""
    from din_b200 import ops
    n, h, w, cin, widths, split_col, norelu = shape
    g = torch.Generator(device="cpu").manual_seed(7)
    x = torch.randn(n, h, w, cin, generator=g).to(cuda).half()
    co = sum(widths)
    wt = (torch.randn(co, cin, 1, 1, generator=g) * (2.0 / cin) ** 0.5).to(cuda)
    bias = torch.randn(co, generator=g).to(cuda)
    pool_bias = bias[norelu[0]:norelu[1]].clone()
    bias[norelu[0]:norelu[1]] = 0
    wp = ops.pack_conv_weight(wt)
    out = torch.full((n, h, w, split_col + 40), 7.0, dtype=torch.float16, device=cuda)      # wider concat buffer
    out2 = torch.full((n, h, w, co - split_col + 8), 7.0, dtype=torch.float16, device=cuda)
    ops.conv2d_branches_nhwc(x, wp, bias, out, out2, split_col=split_col, norelu=norelu, y_c_offset=8)
    torch.cuda.synchronize()
    z = _ref_conv(x, _unpack(wp, cin), bias, 1, (0, 0), False)
    ref = F.relu(z)
    ref[..., norelu[0]:norelu[1]] = z[..., norelu[0]:norelu[1]]
    scale = ref.abs().max().item()
    assert (out[..., :8] == 7).all() and (out[..., 8 + split_col:] == 7).all() and (out2[..., co - split_col:] == 7).all()
    assert (out[..., 8:8 + split_col].float() - ref[..., :split_col]).abs().max().item() <= 2e-3 * scale
    assert (out2[..., :co - split_col].float() - ref[..., split_col:]).abs().max().item() <= 2e-3 * scale
    # the pool branch: 1x1 first, pool + bias + ReLU after
    c = norelu[1] - norelu[0]
    tail = torch.zeros((n, h, w, c + 16), dtype=torch.float16, device=cuda)
    ops.avgpool3_bias_relu_nhwc(out2, pool_bias, tail, c=c, x_c_offset=norelu[0] - split_col, y_c_offset=16)
    torch.cuda.synchronize()
    xp = F.avg_pool2d(x.float().permute(0, 3, 1, 2), 3, 1, 1).permute(0, 2, 3, 1).contiguous().half()
    want = _ref_conv(xp, _unpack(wp, cin)[norelu[0]:norelu[1]], pool_bias, 1, (0, 0), True)
    assert (tail[..., :16] == 0).all()
    assert (tail[..., 16:].float() - want).abs().max().item() <= 3e-3 * want.abs().max().item()


def test_inception_merged_branch_heads_match_the_per_branch_plan(cuda, monkeypatch):
    """Inv3Plan with the branch heads merged (default) vs one launch per branch convolution (DIN_INV3_MERGE=0): the
    multiscale maps agree to fp16 rounding (only the pool branches differ: 1x1 and average pool swapped)."""
    import din_oracle as O
    from din_b200 import inception
    pc = O.PathConfig(backbone="inv3", image_size=(139, 203), out_size=O.backbone_out_size("inv3", 139, 203),
                      emb_features=1056, num_frames=2, num_boxes=4, lite_dim=None)
    sd = {k: v.to(cuda) for k, v in O.make_state_dict(pc, seed=3).items() if k.startswith("backbone.")}
    images = (torch.rand(2, 3, 139, 203, generator=torch.Generator().manual_seed(5)) * 255).to(cuda)
    maps = []
    for merge in (True, False):
        monkeypatch.setattr(inception, "MERGE_1X1", merge)
        maps.append(inception.Inv3Plan(sd)(images).float())
    torch.cuda.synchronize()
    a, b = maps
    assert torch.isfinite(a).all()
    assert (a - b).abs().max().item() <= 4e-3 * b.abs().max().item(), ((a - b).abs().max().item(), b.abs().max().item())


@pytest.mark.parametrize("u8", [False, True], ids=["f32", "u8"])
@pytest.mark.parametrize("n,h,w", [(2, 70, 160), (3, 73, 48), (1, 37, 16), (1, 720, 1280), (2, 480, 720), (5, 96, 272)],
                         ids=str)
def test_resnet_stem_with_fused_maxpool_is_bit_identical(cuda, n, h, w, u8):
    """din_stem7x7_pool_nhwc_f16 (conv1 + folded bn1 + relu + maxpool in one launch) == din_stem_conv_* followed by
    din_maxpool2d_nhwc_f16(3, 2, 1), bit for bit: odd / even row counts, several strips of 63 pooled pixels, several images
    per CTA, image borders."""
    from din_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(h * 31 + w)
    raw = torch.randint(0, 256, (n, h, w, 3), generator=g, dtype=torch.uint8)
    x = raw.to(cuda) if u8 else raw.permute(0, 3, 1, 2).float().contiguous().to(cuda)
    wt = (torch.randn(64, 3, 7, 7, generator=g) * 0.05).to(cuda)
    b = torch.randn(64, generator=g).to(cuda)
    assert ops.stem_pool_supported(x)
    want = ops.maxpool2d_nhwc(ops.stem_conv(x, wt, b, stride=2, pad=3, relu=True, prep=True), 3, 2, 1)
    got = ops.stem_conv_pool(x, wt, b, prep=True)
    torch.cuda.synchronize()
    assert got.shape == want.shape, (tuple(got.shape), tuple(want.shape))
    assert torch.equal(got, want), ((got.float() - want.float()).abs().max().item(),
                                    (got != want).nonzero()[:5].tolist())


def test_resnet_stem_pool_rejects_unaligned_rows(cuda):
    from din_b200 import _lib, ops
    x = torch.zeros(1, 3, 40, 50, device=cuda)                       # 50 * 4 bytes: not a 16-byte multiple
    assert not ops.stem_pool_supported(x)
    with pytest.raises(_lib.DinError, match="16-byte"):
        ops.stem_conv_pool(x, torch.zeros(64, 3, 7, 7, device=cuda), None)
