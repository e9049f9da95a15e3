"""A/B of the stem variants inside ONE process (clocks differ from box to box and over time under the power cap):
interleaved launches of the single-role kernel (DIN_STEM_PIPE=0) and the pipelined one, plus a write-only
reference (fill of the same output size) to see what a pure write stream reaches."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "din-group-activity-recognition-benchmark_b200"))
from din_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
F = 16
img = torch.randint(0, 256, (F, 3, 720, 1280), generator=g).float().to(dev)
u8 = img.permute(0, 2, 3, 1).contiguous().to(torch.uint8)
w0 = (torch.randn(64, 3, 3, 3, generator=g) * 0.2).to(dev)
b0 = torch.randn(64, generator=g).to(dev)


def timed(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3     # us


out_bytes = F * 720 * 1280 * 64 * 2
buf = torch.empty(out_bytes // 2, dtype=torch.float16, device=dev)
src = torch.empty(out_bytes // 2, dtype=torch.float16, device=dev)
res = {}
for rnd in range(3):
    for name, env, x in (("single_f32", "0", img), ("pipe_f32", "1", img), ("single_u8", "0", u8), ("pipe_u8", "1", u8)):
        os.environ["DIN_STEM_PIPE"] = env
        res.setdefault(name, []).append(timed(lambda: ops.stem_conv(x, w0, b0, stride=1, pad=1)))
    res.setdefault("fill_write_only", []).append(timed(lambda: buf.fill_(1.0)))
    res.setdefault("copy_read_write", []).append(timed(lambda: buf.copy_(src)))
for k, v in res.items():
    t = min(v)
    gb = out_bytes / 1e9 * (2 if k.startswith("copy") else 1)
    print(f"{k:18s} best {t:8.1f} us   all {[round(a, 1) for a in v]}   {gb / (t * 1e-6) / 1e3:.2f} TB/s (output bytes"
          f"{' x2' if k.startswith('copy') else ''})")
