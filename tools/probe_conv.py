"""GPU probe: which UMMA-descriptor convention makes the HALO A-operand mode correct?
usage: DIN_CONV_VARIANT=<0..7> python tools/probe_conv.py   (bit0 base_offset, bit1 padded per-row loads, bit2 TAP)"""
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "din-group-activity-recognition-benchmark_b200"))
from din_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
print("variant", os.environ.get("DIN_CONV_VARIANT"))
CASES = [
    (1, 16, 8, 64, 64, (3, 3), (1, 1), False),
    (2, 45, 80, 128, 128, (3, 3), (1, 1), False),
    (1, 90, 160, 256, 256, (3, 3), (1, 1), True),
    (1, 22, 40, 512, 512, (3, 3), (1, 1), False),
    (1, 17, 29, 128, 192, (1, 7), (0, 3), False),
    (1, 17, 29, 128, 192, (7, 1), (3, 0), False),
    (1, 35, 35, 64, 64, (5, 5), (2, 2), False),
    (1, 360, 640, 64, 128, (3, 3), (1, 1), True),
]
g = torch.Generator().manual_seed(0)
for (n, h, w, ci, co, k, pad, pool) in CASES:
    x = torch.randn(n, h, w, ci, generator=g).to(dev).half()
    wt = (torch.randn(co, ci, *k, generator=g) * (2.0 / (ci * k[0] * k[1])) ** 0.5).to(dev)
    b = torch.randn(co, generator=g).to(dev)
    wp = ops.pack_conv_weight(wt)
    try:
        y = ops.conv2d_nhwc(x, wp, b, stride=1, pad=pad, relu=True, pool2=pool)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print("case", (n, h, w, ci, co, k), "EXC", str(e)[:200])
        break
    ref = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), wp[..., :ci].float().permute(0, 3, 1, 2).contiguous(), b, padding=pad))
    if pool:
        ref = F.max_pool2d(ref, 2, 2)
    ref = ref.permute(0, 2, 3, 1)
    err = (y.float() - ref).abs().max().item()
    ok = err <= 2e-3 * ref.abs().max().item()
    # timing
    for _ in range(2):
        ops.conv2d_nhwc(x, wp, b, stride=1, pad=pad, relu=True, pool2=pool)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        ops.conv2d_nhwc(x, wp, b, stride=1, pad=pad, relu=True, pool2=pool)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    fl = 2 * n * h * w * co * ci * k[0] * k[1]
    print(f"case {(n, h, w, ci, co, k)} pool={pool}: err={err:.3e} max|ref|={ref.abs().max().item():.2f} "
          f"{'OK' if ok else 'FAIL'}  {dt * 1e3:.3f} ms  {fl / dt / 1e12:.0f} TFLOP/s")
