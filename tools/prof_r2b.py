"""Representative launches of the kernels added / changed in round 2's second half, for `ncu --set full`:
wgrad (one-CTA vs CTA-pair) at two VGG-16 shapes, the Inception branch-head GEMM + pool tail, the strip upsample and its
adjoint, the pad-0 max-pool backward, the PIL-exact resize."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "din-group-activity-recognition-benchmark_b200"))
from din_b200 import ingest, ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
for (n, h, w, ci, co) in [(20, 90, 160, 512, 512), (20, 180, 320, 256, 256)]:
    x = torch.randn(n, h, w, ci, generator=g).to(dev).half()
    dz = torch.randn(n, h, w, co, generator=g).to(dev).half()
    dw = torch.zeros(co, 3, 3, ci, device=dev)
    for mode in ("0", "1"):
        os.environ["DIN_WGRAD_2CTA"] = mode
        for _ in range(2):
            ops.conv2d_wgrad_nhwc(x, dz, dw, None)
    torch.cuda.synchronize()
    del x, dz
os.environ.pop("DIN_WGRAD_2CTA")
# Mixed_6c branch heads at 720p (27 frames): 768 -> 704 in one GEMM, pool tail
n, h, w = 27, 43, 78
x = torch.randn(n, h, w, 768, generator=g).to(dev).half()
wt = (torch.randn(704, 768, 1, 1, generator=g) * 0.05).to(dev)
wp = ops.pack_conv_weight(wt)
bias = torch.zeros(704, device=dev)
out = torch.empty(n, h, w, 768, dtype=torch.float16, device=dev)
zm = torch.empty(n, h, w, 704, dtype=torch.float16, device=dev)
for _ in range(2):
    ops.conv2d_branches_nhwc(x, wp, bias, out, zm, split_col=192, norelu=(512, 704), y2_c_offset=192)
    ops.avgpool3_bias_relu_nhwc(zm, bias[:192].contiguous(), out, c=192, x_c_offset=512, y_c_offset=576)
# multiscale build and its adjoint
fm = torch.empty(n, 87, 157, 1088, dtype=torch.float16, device=dev)
for _ in range(2):
    ops.upsample_bilinear_nhwc(out, 87, 157, out=fm, c=768, y_c_offset=288)
    ops.upsample_bilinear_bwd_nhwc(fm, 43, 78, c=768, dy_c_offset=288)
y4a = torch.relu(torch.randn(20, 176, 316, 192, generator=g)).to(dev).half()
d = torch.randn(20, 87, 157, 192, generator=g).to(dev).half()
for _ in range(2):
    ops.maxpool3s2_bwd_nhwc(y4a, d, torch.empty_like(y4a), c=192, pad=0)
fr = torch.randint(0, 256, (80, 720, 1280, 3), dtype=torch.uint8, device=dev)
for _ in range(2):
    ingest.resize_u8(fr, (480, 720))
torch.cuda.synchronize()
print("done")
