"""ResNet-18 layer1 (3x3, 64 -> 64 at 180x320, with / without the residual) per 107 frames: CUDA-event times, or -- with
--ncu -- one launch of each for an `ncu --set full` capture."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "din-group-activity-recognition-benchmark_b200"))
import torch
from din_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
n = 107
x = torch.randn(n, 180, 320, 64, generator=g).to(dev).half()
res = torch.randn(n, 180, 320, 64, generator=g).to(dev).half()
wt = (torch.randn(64, 64, 3, 3, generator=g) * (2.0 / 576) ** 0.5).to(dev)
b = torch.randn(64, generator=g).to(dev)
wp = ops.pack_conv_weight(wt)
out = torch.empty_like(x)
reps = 1 if "--ncu" in sys.argv else 10
for nm, r in (("plain", None), ("residual", res)):
    for _ in range(3 if reps > 1 else 1):
        ops.conv2d_nhwc(x, wp, b, stride=1, pad=(1, 1), relu=True, residual=r, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.conv2d_nhwc(x, wp, b, stride=1, pad=(1, 1), relu=True, residual=r, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2 * n * 180 * 320 * 64 * 576
    by = (2 + (1 if r is not None else 0)) * x.numel() * 2
    print(f"{nm}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s  {by / ms / 1e6:.0f} GB/s algorithmic  checksum {out.float().abs().sum().item():.6e}")
