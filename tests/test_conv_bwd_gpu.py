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
