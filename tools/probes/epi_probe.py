import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "din-group-activity-recognition-benchmark_b200"))
import ctypes, torch
from din_b200 import ops, _lib
n, h, w, ci, co, k = (int(a) for a in sys.argv[1:7])
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
x = torch.randn(n, h, w, ci, generator=g).to(dev).half()
wp = ops.pack_conv_weight((torch.randn(co, ci, k, k, generator=g) * 0.05).to(dev))
b = torch.randn(co, generator=g).to(dev)
out = ops.conv2d_nhwc(x, wp, b, stride=1, pad=(k // 2, k // 2), relu=True)
res = torch.randn_like(out) if "--residual" in sys.argv else None
for _ in range(2):
    ops.conv2d_nhwc(x, wp, b, stride=1, pad=(k // 2, k // 2), relu=True, residual=res, out=out)
torch.cuda.synchronize()
L = _lib.load(); L.din_debug_word.restype = ctypes.c_uint
wd = lambda i: int(L.din_debug_word(i))
base = [wd(i) for i in range(40, 46)]
ops.conv2d_nhwc(x, wp, b, stride=1, pad=(k // 2, k // 2), relu=True, residual=res, out=out)
torch.cuda.synchronize()
d = [wd(40 + i) for i in range(6)]
t = 1
d[5] = wd(45)
print(f"{n}x{h}x{w} {ci}->{co}{' +res' if res is not None else ''}: epilogue warp 4 of CTA 0, {d[5]} tiles, cycles per tile: wait tmem_full {d[0] // t}, "
      f"tcgen05.ld wait {d[1] // t}, bias/residual/convert {d[2] // t}, stores {d[3] // t}, fence + arrive {d[4] // t}")
