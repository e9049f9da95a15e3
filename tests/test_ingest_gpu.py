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
