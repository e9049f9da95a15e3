"""Representative launches of the training-step kernels for ncu: wgrad (conv1_2, conv2_2, conv4_2 shapes),
dgrad (= the forward kernel), ReLU/pool backward, stem wgrad, uint8 stem, DIN backward."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "din-group-activity-recognition-benchmark_b200"))
from din_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
REP = 2
CASES = [(4, 720, 1280, 64, 64), (8, 360, 640, 128, 128), (8, 90, 160, 512, 512)]
for (n, h, w, ci, co) in CASES:
    x = torch.randn(n, h, w, ci, generator=g).to(dev).half()
    dz = torch.randn(n, h, w, co, generator=g).to(dev).half()
    dw = torch.zeros(co, 3, 3, ci, device=dev)
    db = torch.zeros(co, device=dev)
    for _ in range(REP):
        ops.conv2d_wgrad_nhwc(x, dz, dw, db)
    y = torch.relu(torch.randn(n, h, w, co, generator=g)).to(dev).half()
    dp = torch.randn(n, h // 2, w // 2, co, generator=g).to(dev).half()
    for _ in range(REP):
        ops.relu_pool_bwd_nhwc(y, dp, True)
    torch.cuda.synchronize()
    del x, dz, y, dp
img = torch.randint(0, 256, (4, 720, 1280, 3), generator=g, dtype=torch.uint8).to(dev)
w0 = (torch.randn(64, 3, 3, 3, generator=g) * 0.2).to(dev)
b0 = torch.randn(64, generator=g).to(dev)
dz = torch.randn(4, 720, 1280, 64, generator=g).to(dev).half()
for _ in range(REP):
    ops.stem_conv(img, w0, b0, stride=1, pad=1)
    ops.stem_wgrad(img, dz, torch.zeros(64, 3, 3, 3, device=dev), torch.zeros(64, device=dev))
# Dynamic Relation / Walk backward at the headline shape (B=8, T=10, N=12, C=128) and at C=1024
for C in (128, 1024):
    x = torch.randn(8, 10, 12, C, generator=g).to(dev)
    w_tap = (torch.randn(9, 27, C, generator=g) * 0.01).to(dev)
    b_cat = (torch.randn(27, generator=g) * 0.3).to(dev)
    dy = torch.randn(8, 10, 12, C, generator=g).to(dev)
    for _ in range(REP):
        ops.dynamic_infer(x, w_tap, b_cat, (3, 3), 1)
        ops.dynamic_infer_bwd(x, w_tap, b_cat, dy, torch.zeros_like(x), (3, 3), 1)
torch.cuda.synchronize()
print("done")
