"""ResNet-18 stem (7x7 stride 2, space-to-depth, warp-specialised): time per 107 frames at 720p under the knobs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "din-group-activity-recognition-benchmark_b200"))
import torch
from din_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
x = torch.randint(0, 256, (107, 3, 720, 1280), generator=g).float().to(dev)
w = (torch.randn(64, 3, 7, 7, generator=g) * 0.05).to(dev)
b = torch.randn(64, generator=g).to(dev)
for _ in range(3):
    y = ops.stem_conv(x, w, b, stride=2, pad=3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    y = ops.stem_conv(x, w, b, stride=2, pad=3)
e1.record()
torch.cuda.synchronize()
print(f"stem alone: {e0.elapsed_time(e1) / 10:.3f} ms per 107 frames, checksum {y.float().abs().sum().item():.6e}")
for _ in range(3):
    z = ops.stem_conv_pool(x, w, b)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    z = ops.stem_conv_pool(x, w, b)
e1.record()
torch.cuda.synchronize()
print(f"stem + pool fused: {e0.elapsed_time(e1) / 10:.3f} ms per 107 frames, checksum {z.float().abs().sum().item():.6e}")
xu = x.permute(0, 2, 3, 1).contiguous().to(torch.uint8)
for fn, nm in ((lambda: ops.stem_conv(xu, w, b, stride=2, pad=3), "stem alone u8"), (lambda: ops.stem_conv_pool(xu, w, b), "fused u8")):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{nm}: {e0.elapsed_time(e1) / 10:.3f} ms per 107 frames")
