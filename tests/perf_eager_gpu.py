"""Context measurement (not a test, not part of bench.py): the reference's own PyTorch path executed by
PyTorch-eager CUDA (cuDNN / cuBLAS) on the same B200 — the 'honest GPU bar' of BASELINE.md §4 — next to this
repo's kernels, on the VGG-16 backbone of the headline workload (16 frames, 720x1280).

usage (GPU box):  python tests/perf_eager_gpu.py
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "din-group-activity-recognition-benchmark_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import din_oracle as O  # noqa: E402


def timed(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device("cuda:0")
    frames = 16
    gflop = 563.75 * frames
    pc = O.PathConfig()
    bb = O.build_backbone("vgg16")
    sd = O.make_state_dict(pc, seed=0, backbone=bb)
    O.load_backbone(bb, sd)
    x = torch.randint(0, 256, (frames, 3, 720, 1280), device=dev).float()
    res = {}
    with torch.no_grad():
        m = bb.to(dev).eval()
        torch.backends.cudnn.benchmark = True
        for name, tf32 in (("eager fp32 (TF32 off)", False), ("eager fp32 (TF32 on)", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            ms = timed(lambda: m(O.prep_images(x)))
            res[name] = {"ms": ms, "TFLOP/s": gflop / ms}
        mh = m.half().to(memory_format=torch.channels_last)
        xh = O.prep_images(x).half().contiguous(memory_format=torch.channels_last)
        ms = timed(lambda: mh(xh))
        res["eager fp16 channels_last (cuDNN, prep excluded)"] = {"ms": ms, "TFLOP/s": gflop / ms}
    # this repo
    from din_b200.engine import VGG16Plan
    sdd = {k: v.to(dev) for k, v in sd.items()}
    plan = VGG16Plan(sdd)
    ms = timed(lambda: plan(x))
    res["din_b200 (tcgen05 conv + TC stem, prep fused)"] = {"ms": ms, "TFLOP/s": gflop / ms}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
