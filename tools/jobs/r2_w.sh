#!/bin/bash
# ingest: resize kernel vs oracle / Pillow, nvJPEG decode; bn_fold_grads unit test; decode throughput
O=gpurun_out/r2w; mkdir -p $O
timeout 900 python -m pytest tests/test_ingest_gpu.py tests/test_conv_bwd_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -rA -s -k "resize or jpeg or decoded or bn_fold" > $O/pytest.log 2>&1; echo "rc=$?"
grep -E "passed|failed|error" $O/pytest.log | tail -3; grep -E "^FAILED|^ERROR|^E  |\[jpeg" $O/pytest.log | head -40
timeout 300 python - <<'PY' 2>&1 | tail -12
import io, sys, time
sys.path.insert(0, 'din-group-activity-recognition-benchmark_b200')
import numpy as np, torch
from PIL import Image
from din_b200 import ingest
rng = np.random.default_rng(0)
def jpeg(h, w):
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([(xx * 255 // (w - 1)), (yy * 255 // (h - 1)), ((xx + yy) * 255 // (h + w - 2))], -1)
    img = (img + 40 * np.sin(xx / 9.0)[..., None] + rng.integers(-25, 26, size=(h, w, 3))).clip(0, 255).astype(np.uint8)
    b = io.BytesIO(); Image.fromarray(img).save(b, format='JPEG', quality=90, subsampling=2); return b.getvalue()
for (src, dst, n) in (((720, 1280), (720, 1280), 80), ((720, 1280), (480, 720), 80), ((480, 640), (480, 720), 80)):
    js = [jpeg(*src) for _ in range(4)] * (n // 4)
    for thr in (1, 8, 16):
        ingest.decode_resize(js, dst, cpu_threads=thr)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3): ingest.decode_resize(js, dst, cpu_threads=thr)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
        print(f"decode+resize {src}->{dst} n={n} threads={thr}: {dt*1e3:.1f} ms  {n/dt:.0f} frames/s  ({sum(map(len,js))/n/1e3:.0f} kB/frame)")
    fr = torch.randint(0, 256, (n,) + src + (3,), dtype=torch.uint8, device='cuda')
    if src != dst:
        ingest.resize_u8(fr, dst); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ingest.resize_u8(fr, dst)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        byt = n * 3 * (src[0]*src[1] + 2*src[0]*dst[1] + dst[0]*dst[1])
        print(f"resize only: {ms:.3f} ms per {n} frames = {byt/ms/1e6:.0f} GB/s algorithmic (incl. allocation)")
t0=time.perf_counter()
for j in js[:20]: np.array(Image.open(io.BytesIO(j)).resize((720, 480), Image.BILINEAR))
print(f"PIL decode+resize 1 thread: {20/(time.perf_counter()-t0):.0f} frames/s")
PY
