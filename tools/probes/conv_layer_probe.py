"""One stride-1 convolution layer through din_conv2d_nhwc_f16: CUDA-event time, or (--ncu) one launch for `ncu --set full`.
usage: conv_layer_probe.py n h w c_in c_out k [--residual] [--ncu]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "din-group-activity-recognition-benchmark_b200"))
import torch
from din_b200 import ops
n, h, w, ci, co, k = (int(a) for a in sys.argv[1:7])
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
x = torch.randn(n, h, w, ci, generator=g).to(dev).half()
wt = (torch.randn(co, ci, k, k, generator=g) * (2.0 / (ci * k * k)) ** 0.5).to(dev)
b = torch.randn(co, generator=g).to(dev)
wp = ops.pack_conv_weight(wt)
pad = (k // 2, k // 2)
out = ops.conv2d_nhwc(x, wp, b, stride=1, pad=pad, relu=True)
res = torch.randn_like(out) if "--residual" in sys.argv else None
reps = 1 if "--ncu" in sys.argv else 10
for _ in range(3 if reps > 1 else 1):
    ops.conv2d_nhwc(x, wp, b, stride=1, pad=pad, relu=True, residual=res, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ops.conv2d_nhwc(x, wp, b, stride=1, pad=pad, relu=True, residual=res, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
fl = 2 * out.numel() * ci * k * k
by = (x.numel() + out.numel() * (2 if res is not None else 1)) * 2
print(f"{n}x{h}x{w} {ci}->{co} k{k}{' +res' if res is not None else ''}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s  {by / ms / 1e6:.0f} GB/s algorithmic")
