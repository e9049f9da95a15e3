"""uint8 NHWC ingest: the stem fed with the decoded frame as the loader holds it before
`img.transpose(2,0,1)` / `.float()` (volleyball.py:239-243,270) must give the SAME BITS as the fp32 NCHW
path on equal pixel values (uint8 -> fp32 is exact), at the kernel and at the model level."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("geom", [(64, 3, 1, 1), (64, 7, 2, 3), (32, 3, 2, 0)], ids=["vgg", "res18", "inv3"])
@pytest.mark.parametrize("hw", [(70, 150), (33, 257), (720, 1280)], ids=["70x150", "33x257", "720p"])
def test_stem_u8_bit_identical_to_f32(cuda, geom, hw):
    from din_b200 import ops
    co, k, s, p = geom
    g = torch.Generator(device="cpu").manual_seed(11)
    n = 2 if hw[0] < 700 else 1
    u8 = torch.randint(0, 256, (n, hw[0], hw[1], 3), generator=g, dtype=torch.uint8).to(cuda)
    f32 = u8.permute(0, 3, 1, 2).float().contiguous()
    wt = (torch.randn(co, 3, k, k, generator=g) * 0.1).to(cuda)
    b = torch.randn(co, generator=g).to(cuda)
    for prep in (True, False):
        ya = ops.stem_conv(f32, wt, b, stride=s, pad=p, relu=True, prep=prep)
        yb = ops.stem_conv(u8, wt, b, stride=s, pad=p, relu=True, prep=prep)
        torch.cuda.synchronize()
        assert ya.shape == yb.shape
        assert torch.equal(ya, yb), (geom, hw, prep, (ya.float() - yb.float()).abs().max().item())


def test_stem_u8_rejects_other_geometries(cuda):
    from din_b200 import ops
    from din_b200._lib import DinError
    u8 = torch.zeros((1, 16, 16, 3), dtype=torch.uint8, device=cuda)
    with pytest.raises(DinError, match="only the three backbone stems"):
        ops.stem_conv(u8, torch.zeros(64, 3, 5, 5, device=cuda), torch.zeros(64, device=cuda), stride=1, pad=2)


@pytest.mark.parametrize("backbone,hw", [("vgg16", (96, 160)), ("res18", (96, 160)), ("inv3", (139, 203))])
def test_model_u8_frames_equal_f32_frames(cuda, backbone, hw):
    """Dynamic_volleyball((uint8 [B,T,H,W,3], boxes)) == Dynamic_volleyball((fp32 [B,T,3,H,W], boxes)), bit for bit."""
    import din_oracle as O
    import infer_model as IM
    from config import Config
    kw = dict(emb_features=1056, lite_dim=None) if backbone == "inv3" else {}
    pc = O.PathConfig(backbone=backbone, image_size=hw, out_size=O.backbone_out_size(backbone, *hw), num_frames=3,
                      num_boxes=4, **kw)
    bb = O.build_backbone(backbone)
    sd = O.make_state_dict(pc, seed=2, backbone=bb)
    images, boxes = O.make_inputs(pc, 2, seed=2)                       # fp32 [B,T,3,H,W], integer-valued
    cfg = Config("volleyball")
    cfg.log_path = None
    for k in ("backbone", "image_size", "out_size", "emb_features", "num_frames", "num_boxes", "lite_dim",
              "ST_kernel_size", "scale_factor", "beta_factor", "hierarchical_inference", "num_DIM"):
        setattr(cfg, k, getattr(pc, k))
    cfg.sampling_ratio = list(pc.sampling_ratio)
    model = IM.Dynamic_volleyball(cfg)
    model.load_state_dict(sd, strict=True)
    model = model.to(cuda).eval()
    u8 = images.permute(0, 1, 3, 4, 2).contiguous().to(torch.uint8)   # [B,T,H,W,3]
    assert torch.equal(u8.float().permute(0, 1, 4, 2, 3), images)
    with torch.no_grad():
        a = model((images.to(cuda), boxes.to(cuda)))["activities"].clone()
        b = model((u8.to(cuda), boxes.to(cuda)))["activities"].clone()
    torch.cuda.synchronize()
    assert torch.equal(a, b), (a - b).abs().max().item()


# ---- JPEG decode + PIL-exact resize on the device (din_b200/ingest.py, csrc/ingest.cu) ---------------------------------
RESIZE_SHAPES = [((48, 64), (36, 54)), ((30, 40), (48, 72)), ((45, 80), (45, 64)), ((37, 53), (20, 53)), ((7, 5), (3, 11)),
                 ((48, 64), (48, 64))]


@pytest.mark.parametrize("src,dst", RESIZE_SHAPES, ids=str)
def test_resize_kernel_is_bit_identical_to_the_oracle(cuda, src, dst):
    import numpy as np
    import pil_resize_oracle as R
    from din_b200 import ingest
    rng = np.random.default_rng(src[0] * 100 + dst[1])
    imgs = rng.integers(0, 256, size=(3,) + src + (3,), dtype=np.uint8)
    imgs[:, : src[0] // 3] = 255
    imgs[:, -(src[0] // 4):, : src[1] // 2] = 0
    got = ingest.resize_u8(torch.from_numpy(imgs).to(cuda), dst).cpu().numpy()
    for i in range(3):
        assert np.array_equal(got[i], R.resize_bilinear_u8(imgs[i], dst)), (src, dst, i)


def test_resize_kernel_at_dataset_sizes_equals_pillow(cuda):
    """Full-size frames: 720 x 1280 -> Collective's 480 x 720 (scripts/train_collective_stage2_dynamic.py:17), and
    480 x 640 -> 480 x 720 (Collective's own frame size), against Pillow itself."""
    import numpy as np
    Image = pytest.importorskip("PIL.Image")
    from din_b200 import ingest
    rng = np.random.default_rng(0)
    for src, dst in (((720, 1280), (480, 720)), ((480, 640), (480, 720)), ((1080, 1920), (720, 1280))):
        base = rng.integers(0, 256, size=(src[0] // 8 + 1, src[1] // 8 + 1, 3), dtype=np.uint8)
        img = np.kron(base, np.ones((8, 8, 1), dtype=np.uint8))[: src[0], : src[1]]          # blocky: edges everywhere
        img = (img.astype(np.int16) + rng.integers(-6, 7, size=img.shape)).clip(0, 255).astype(np.uint8)
        want = np.array(Image.fromarray(img).resize((dst[1], dst[0]), Image.BILINEAR))
        got = ingest.resize_u8(torch.from_numpy(img[None]).to(cuda), dst)[0].cpu().numpy()
        assert np.array_equal(got, want), (src, dst, int(np.abs(got.astype(int) - want.astype(int)).max()))


def _synthetic_jpeg(rng, h, w, subsampling, quality=92):
    import io
    import numpy as np
    from PIL import Image
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([(xx * 255 // max(1, w - 1)), (yy * 255 // max(1, h - 1)), ((xx + yy) * 255 // max(1, h + w - 2))], -1)
    img = (img + 40 * np.sin(xx / 9.0)[..., None] + rng.integers(-10, 11, size=(h, w, 3))).clip(0, 255).astype(np.uint8)
    img[h // 4: h // 2, w // 3: w // 2] = (220, 30, 40)                                  # a sharp coloured patch
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format="JPEG", quality=quality, subsampling=subsampling)
    data = buf.getvalue()
    return data, np.array(Image.open(io.BytesIO(data)).convert("RGB"))


@pytest.mark.parametrize("subsampling", [0, 2], ids=["444", "420"])
def test_jpeg_decode_resize_matches_the_loader(cuda, subsampling):
    """nvJPEG decode + the exact resize vs Image.open + resize (volleyball.py:237-240).  The two JPEG decoders differ in
    their IDCT / colour-conversion rounding (4:4:4: +-1 on half of the samples) and in chroma upsampling at sharp colour edges (4:2:0); the resize adds
    nothing: decode_resize == resize_u8(decode at native size) bit for bit."""
    import numpy as np
    Image = pytest.importorskip("PIL.Image")
    from din_b200 import ingest
    rng = np.random.default_rng(5)
    sizes = [(96, 160), (120, 144), (96, 160)]                       # the first and the last already have the target size
    target = (96, 160)
    jpegs, refs = zip(*[_synthetic_jpeg(rng, h, w, subsampling) for h, w in sizes])
    assert [ingest.jpeg_size(j) for j in jpegs] == sizes
    out = ingest.decode_resize(jpegs, target, device=cuda)
    assert out.shape == (3,) + target + (3,) and out.dtype == torch.uint8
    got = out.cpu().numpy()
    for i, (h, w) in enumerate(sizes):
        want = np.array(Image.fromarray(refs[i]).resize((target[1], target[0]), Image.BILINEAR))
        d = np.abs(got[i].astype(int) - want.astype(int))
        print(f"\n[jpeg {['444', '', '420'][subsampling]}] frame {i} {h}x{w}: max|Δ| {d.max()}, mean|Δ| {d.mean():.4f}, "
              f"share > 1: {(d > 1).mean():.5f}")
        if subsampling == 0:
            # measured: half of the samples differ by one level (the two decoders round the YCbCr -> RGB conversion
            # differently), 1 % by two, none by more than three
            assert d.max() <= 4 and d.mean() <= 0.75 and (d > 1).mean() <= 0.03, (d.max(), d.mean())
        else:
            assert d.mean() <= 1.5 and (d > 8).mean() <= 0.02, (d.max(), d.mean(), (d > 8).mean())
    # the resize is exact: decoding at native size and resizing separately gives the same bytes
    native = ingest.decode_resize([jpegs[1]], sizes[1], device=cuda)
    assert torch.equal(ingest.resize_u8(native, target)[0], out[1])
    with pytest.raises(Exception, match="JPEG"):
        ingest.decode_resize([b"not a jpeg stream at all" * 4], target, device=cuda)


def test_decoded_frames_feed_the_uint8_stem(cuda):
    """decode_resize's tensor is what the models take as uint8 frames: logits equal those of the same pixels passed as the
    reference loader's fp32 NCHW tensor."""
    import din_oracle as O
    import infer_model as IM
    from config import Config
    from din_b200 import ingest
    import numpy as np
    rng = np.random.default_rng(8)
    pc = O.PathConfig(backbone="vgg16", image_size=(96, 160), out_size=O.backbone_out_size("vgg16", 96, 160), num_frames=3,
                      num_boxes=4)
    jpegs = [_synthetic_jpeg(rng, 96 + 24 * (i % 2), 160, 2)[0] for i in range(6)]        # every other frame needs resizing
    frames = ingest.decode_resize(jpegs, (96, 160), device=cuda).view(2, 3, 96, 160, 3)
    cfg = Config("volleyball")
    cfg.log_path = None
    for k in ("backbone", "image_size", "out_size", "emb_features", "num_frames", "num_boxes", "crop_size",
              "num_features_boxes", "num_activities", "lite_dim", "ST_kernel_size", "scale_factor", "beta_factor",
              "hierarchical_inference", "num_DIM"):
        setattr(cfg, k, getattr(pc, k))
    cfg.sampling_ratio = list(pc.sampling_ratio)
    model = IM.Dynamic_volleyball(cfg)
    model.load_state_dict(O.make_state_dict(pc, seed=1), strict=True)
    model = model.to(cuda).eval()
    _, boxes = O.make_inputs(pc, 2, seed=1)
    with torch.no_grad():
        a = model((frames, boxes.to(cuda)))["activities"]
        b = model((frames.permute(0, 1, 4, 2, 3).float().contiguous(), boxes.to(cuda)))["activities"]
    assert torch.isfinite(a).all() and torch.equal(a, b)


@pytest.mark.parametrize("u8", [False, True], ids=["f32", "u8"])
def test_host_memory_frames_are_streamed_chunk_by_chunk(cuda, u8, monkeypatch):
    """model((frames, boxes)) with the frames in pinned HOST memory (eval mode): the engine copies them chunk by chunk on
    its copy stream under the backbone's kernels; logits are bit-identical to the device-tensor call, call after call
    (three rotating staging buffers, ragged last chunk)."""
    import din_oracle as O
    import infer_model as IM
    from config import Config
    pc = O.PathConfig(backbone="vgg16", image_size=(96, 160), out_size=O.backbone_out_size("vgg16", 96, 160), num_frames=5,
                      num_boxes=4)
    cfg = Config("volleyball")
    cfg.log_path = None
    for k in ("backbone", "image_size", "out_size", "emb_features", "num_frames", "num_boxes", "crop_size",
              "num_features_boxes", "num_activities", "lite_dim", "ST_kernel_size", "scale_factor", "beta_factor",
              "hierarchical_inference", "num_DIM"):
        setattr(cfg, k, getattr(pc, k))
    cfg.sampling_ratio = list(pc.sampling_ratio)
    model = IM.Dynamic_volleyball(cfg)
    model.load_state_dict(O.make_state_dict(pc, seed=2), strict=True)
    model = model.to(cuda).eval()
    model.engine().frames_per_chunk = 4                      # 15 frames -> chunks of 4, 4, 4, 3
    outs = []
    for seed in (1, 2, 3):
        images, boxes = O.make_inputs(pc, 3, seed=seed)
        if u8:
            images = images.permute(0, 1, 3, 4, 2).contiguous().to(torch.uint8)
        with torch.no_grad():
            want = model((images.to(cuda), boxes.to(cuda)))["activities"]
            got = model((images.pin_memory(), boxes.pin_memory()))["activities"]
            got_pageable = model((images, boxes))["activities"]
        assert torch.equal(got, want) and torch.equal(got_pageable, want), seed
        outs.append(want)
    assert not torch.equal(outs[0], outs[1])
    model.train()
    with pytest.raises(RuntimeError, match="eval mode"):
        model((images, boxes))
