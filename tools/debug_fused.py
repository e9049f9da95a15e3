"""Smallest case of the fused conv1_1 + conv1_2 kernel with the wait-timeout recorder on (DIN_FUSED_DEBUG=1)."""
import ctypes as C
import os
import sys

os.environ["DIN_FUSED_DEBUG"] = "1"
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "din-group-activity-recognition-benchmark_b200"))
from din_b200 import _lib, ops  # noqa: E402

n, h, w = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (1, 16, 16)))
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
img = torch.randint(0, 256, (n, 3, h, w), generator=g).float().to(dev)
w1 = (torch.randn(64, 3, 3, 3, generator=g) * 0.3).to(dev)
b1 = (torch.randn(64, generator=g) * 0.2).to(dev)
w2 = (torch.randn(64, 64, 3, 3, generator=g) * (2.0 / 576) ** 0.5).to(dev)
b2 = (torch.randn(64, generator=g) * 0.1).to(dev)
w2p = ops.pack_conv_weight(w2)
want = ops.conv2d_nhwc(ops.stem_conv(img, w1, b1, stride=1, pad=1), w2p, b2, stride=1, pad=(1, 1), relu=True, pool2=False)
torch.cuda.synchronize()
lib = _lib.load()
NAMES = {1: "patch_empty", 2: "col_full", 3: "stem_free", 4: "tmem_empty", 5: "a_full", 6: "b_full", 7: "stem_done(build)",
         8: "patch_full", 9: "stem_done(drain)", 10: "a_empty"}
try:
    got = ops.stem_conv_pair(img, w1, b1, w2p, b2, relu=True, pool2=False)
    torch.cuda.synchronize()
    d = (got.float() - want.float()).abs()
    print(f"OK {n}x{h}x{w}: {int((d > 0).sum())} of {d.numel()} differ, max {d.max().item():.3e}, max|want| {want.float().abs().max().item():.3f}")
    if int((d > 0).sum()):
        bad = (d > 0).nonzero()
        print("first differing (img, y, x, c):", bad[:8].tolist())
        ys, xs = bad[:, 1].unique().tolist(), bad[:, 2].unique().tolist()
        print("rows with differences:", ys[:40], "cols:", xs[:40])
except Exception as e:  # noqa: BLE001
    print("FAILED:", str(e).splitlines()[0])
    w0 = lib.din_debug_word(0)
    print(f"first timeout: code {w0 >> 24} ({NAMES.get(w0 >> 24, '?')}), block {(w0 >> 12) & 0xFFF}, thread {w0 & 0xFFF}")
    for c in range(1, 11):
        v = lib.din_debug_word(1 + c)
        if v:
            print(f"  code {c} {NAMES[c]}: parity {v >> 31}, block {(v >> 12) & 0x7FFFF}, thread {v & 0xFFF}")
