"""RoIAlign and the fused Dynamic Relation / Dynamic Walk kernel at the bench shapes, for ncu:
  headline   VGG-16 map [80, 22, 40, 512] fp16, 960 boxes -> crops; x [8, 10, 12, 128] (lite) through the 3x3 field
  Inv3       multiscale map [80, 87, 157, 1088] fp16 (1056 real channels), 960 boxes; x [8, 10, 12, 1024]
usage: ncu --set full -k regex:"roi_align_kernel|dynamic_infer_kernel" ... python tools/prof_head.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "din-group-activity-recognition-benchmark_b200"))
from din_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
B, T, N = 8, 10, 12
for (oh, ow, d, c) in ((22, 40, 512, 128), (87, 157, 1088, 1024)):
    fm = torch.randn(B * T, oh, ow, d, generator=g).half().to(dev)
    cx, cy = torch.rand(B * T * N, generator=g) * ow, torch.rand(B * T * N, generator=g) * oh
    bw, bh = 1 + 3 * torch.rand(B * T * N, generator=g), 2 + 5 * torch.rand(B * T * N, generator=g)
    boxes = torch.stack((cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2), dim=-1).to(dev)
    idx = torch.arange(B * T, dtype=torch.int32).repeat_interleave(N).to(dev)
    x = torch.randn(B, T, N, c, generator=g).to(dev)
    p_w = (torch.randn(18, c, 3, 3, generator=g) * 0.01).to(dev)
    s_w = (torch.randn(9, c, 3, 3, generator=g) * 0.01).to(dev)
    p_b, s_b = (torch.randn(18, generator=g) * 0.3).to(dev), (torch.randn(9, generator=g) * 0.3).to(dev)
    w_tap, b_cat = ops.pack_din_weights(p_w, p_b, s_w, s_b)
    for _ in range(3):
        crops = ops.roi_align_nhwc(fm, boxes, idx, 5, 5, d=d)
        y = ops.dynamic_infer(x, w_tap, b_cat, (3, 3), 1, scale_factor=True)
    torch.cuda.synchronize()
    print(f"map {oh}x{ow}x{d}: roi_align algorithmic bytes = map {fm.numel() * 2 / 1e6:.1f} MB + crops {crops.numel() * 2 / 1e6:.1f} MB; "
          f"dynamic_infer: x + y = {2 * x.numel() * 4 / 1e6:.2f} MB + weights {w_tap.numel() * 4 / 1e6:.2f} MB")
