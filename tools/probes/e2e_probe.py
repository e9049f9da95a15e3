"""Where does a host-frames forward spend its time?  host-side duration of model(...) (does anything block?) and the
device-side step time, for host (pinned) vs device inputs."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
pc_kw, B, _ = bench.WORKLOADS[os.environ.get("W", bench.DEFAULT_WORKLOAD)]
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import din_oracle as O
pc = O.PathConfig(**pc_kw)
dev = torch.device("cuda:0")
model, sd, bb = bench.build_model(pc, dev)
images, boxes = O.make_inputs(pc, B, seed=0)[:2]
img_h, box_h = images.pin_memory(), boxes.pin_memory()
img_d, box_d = images.to(dev), boxes.to(dev)
print("pinned:", img_h.is_pinned(), img_h.reshape((-1,) + tuple(img_h.shape[2:]))[3:9].is_pinned())
for name, args in (("device", (img_d, box_d)), ("host", (img_h, box_h))):
    with torch.no_grad():
        for _ in range(2):
            model(args)
        torch.cuda.synchronize()
        host_ms, t0 = [], time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            h0 = time.perf_counter()
            model(args)
            host_ms.append((time.perf_counter() - h0) * 1e3)
        e1.record()
        torch.cuda.synchronize()
        print(f"{name}: device {e0.elapsed_time(e1) / 8:.2f} ms/step, host call {sum(host_ms) / 8:.2f} ms/step "
              f"(min {min(host_ms):.2f}, max {max(host_ms):.2f})")
# which kernels slow down while the copies run?
from din_b200 import ops
per = {}
for name, args in (("device", (img_d, box_d)), ("host", (img_h, box_h))):
    with torch.no_grad():
        model(args); torch.cuda.synchronize()
        ops.RECORDER = []
        for _ in range(4):
            model(args)
        torch.cuda.synchronize()
        rec, ops.RECORDER = ops.RECORDER, None
    d = {}
    for (n, f, b, a, z) in rec:
        d[n] = d.get(n, 0.0) + a.elapsed_time(z) / 4
    per[name] = d
for n in per["device"]:
    print(f"{n:50s} device {per['device'][n]:7.3f}  host {per['host'].get(n, 0):7.3f}  x{per['host'].get(n, 0) / max(per['device'][n], 1e-9):.3f}")
print("sum", sum(per["device"].values()), sum(per["host"].values()))
# raw copy throughput: whole tensor vs 3 chunks on a side stream
st = torch.cuda.Stream()
flat = img_h.reshape((-1,) + tuple(img_h.shape[2:]))
buf = torch.empty_like(img_d).reshape(flat.shape)
for label, parts in (("1 copy", [(0, flat.shape[0])]), ("3 copies", [(0, 27), (27, 54), (54, flat.shape[0])])):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(st):
        for a, b in parts:
            buf[a:b].copy_(flat[a:b], non_blocking=True)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{label}: enqueue {1e3 * (t1 - t0):.2f} ms, total {1e3 * (t2 - t0):.2f} ms = {flat.numel() * 4 / (t2 - t0) / 1e9:.1f} GB/s")
