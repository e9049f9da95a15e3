#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu_27.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -4 gpurun_out/pytest_gpu_27.log
timeout 600 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_27.json 2> gpurun_out/bench_27.err
echo "bench rc=$?"; python - <<PY
import json
d=json.load(open('gpurun_out/bench_27.json')); r=d['roofline']
print('clips/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],2), 'convTF', round(r['achieved']), r['other_kernels_ms'], d['clocks'])
print('train_step', json.dumps(d.get('train_step')))
PY
tail -3 gpurun_out/bench_27.err
