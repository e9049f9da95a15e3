#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_18.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -6 gpurun_out/pytest_gpu_18.log
timeout 600 python bench.py > gpurun_out/bench_18.json 2> gpurun_out/bench_18.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_18.json')); r=d['roofline']
print('clips/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],2), 'convTF', round(r['achieved']), r['other_kernels_ms'], d['clocks'], 'cpu', d['cpu_baseline']['value'])
PY
tail -3 gpurun_out/bench_18.err
