#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_backward_gpu.py tests/test_ingest_gpu.py tests/test_conv_gpu.py -m gpu -q -s -k "full_training or ingest or u8 or stem" > gpurun_out/pytest_19.log 2>&1
echo "rc=$?"; grep -E "passed|failed|^FAILED|^E  |\[full step\] loss" gpurun_out/pytest_19.log | cut -c1-200 | head -20
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_19.json 2> gpurun_out/bench_19.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_19.json')); r=d['roofline']
print('clips/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],2), 'convTF', round(r['achieved']), r['other_kernels_ms'], d['clocks'])
print('train_step', json.dumps(d.get('train_step')))
PY
tail -3 gpurun_out/bench_19.err
