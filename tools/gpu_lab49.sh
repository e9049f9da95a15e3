#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_bwd_gpu.py -m gpu -q -k "maxpool3s2 or bn_gamma" > gpurun_out/pytest_49.log 2>&1
echo "rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest_49.log | cut -c1-250 | head -20
python - <<'PY'
import sys, torch
sys.path.insert(0, "din-group-activity-recognition-benchmark_b200")
from din_b200 import ops
dev = torch.device("cuda:0")
x = torch.relu(torch.randn(20, 360, 640, 64, device=dev)).half()
dy = torch.randn(20, 180, 320, 64, device=dev).half()
dz = torch.randn(20, 360, 640, 64, device=dev).half()
gamma = torch.rand(64, device=dev) + 0.5; beta = torch.randn(64, device=dev); dg = torch.zeros(64, device=dev)
def t(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
ms = t(lambda: ops.maxpool3s2_relu_bwd_nhwc(x, dy)); by = 2 * (2 * x.numel() + dy.numel())
print(f"maxpool3s2_bwd {ms:.3f} ms  {by/ms/1e6:.0f} GB/s")
ms = t(lambda: ops.bn_gamma_grad(dz, x, gamma, beta, dg)); by = 4 * x.numel()
print(f"bn_gamma (no sub) {ms:.3f} ms  {by/ms/1e6:.0f} GB/s")
ms = t(lambda: ops.bn_gamma_grad(dz, x, gamma, beta, dg, sub=dz)); by = 4 * x.numel()
print(f"bn_gamma (sub aliased) {ms:.3f} ms")
x5 = torch.relu(torch.randn(20, 23, 40, 512, device=dev)).half(); d5 = torch.randn_like(x5)
g5 = torch.rand(512, device=dev) + 0.5; b5 = torch.randn(512, device=dev); dg5 = torch.zeros(512, device=dev)
ms = t(lambda: ops.bn_gamma_grad(d5, x5, g5, b5, dg5)); print(f"bn_gamma layer4 {ms:.3f} ms  {4*x5.numel()/ms/1e6:.0f} GB/s")
PY
