"""Representative launches for ncu: VGG stem, conv1_2+pool, conv2_1, conv4_2."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "din-group-activity-recognition-benchmark_b200"))
from din_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
img = torch.randint(0, 256, (4, 3, 720, 1280), generator=g).float().to(dev)
w0 = (torch.randn(64, 3, 3, 3, generator=g) * 0.2).to(dev)
b0 = torch.randn(64, generator=g).to(dev)
for _ in range(3):
    y0 = ops.stem_conv(img, w0, b0, stride=1, pad=1)
torch.cuda.synchronize()
CASES = [(4, 720, 1280, 64, 64, True), (8, 360, 640, 64, 128, False), (8, 90, 160, 512, 512, False)]
for (n, h, w, ci, co, pool) in CASES:
    x = torch.randn(n, h, w, ci, generator=g).to(dev).half()
    wt = (torch.randn(co, ci, 3, 3, generator=g) * (2.0 / (ci * 9)) ** 0.5).to(dev)
    b = torch.randn(co, generator=g).to(dev)
    wp = ops.pack_conv_weight(wt)
    for _ in range(3):
        y = ops.conv2d_nhwc(x, wp, b, stride=1, pad=(1, 1), relu=True, pool2=pool)
    torch.cuda.synchronize()
