#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do
DIN_CONV_2CTA=$v timeout 600 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_31_2cta$v.json 2> gpurun_out/bench_31.err
echo "bench 2cta=$v rc=$?"; python - <<PY
import json
d=json.load(open('gpurun_out/bench_31_2cta$v.json')); r=d['roofline']
print('clips/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],2), 'convTF', round(r['achieved']), 'frac', round(r['frac'],3), d['clocks'])
print({k:v for k,v in r['per_layer_tflops'].items() if '64->' in k or '128->128' in k})
print('train_step', d['train_step']['ms_per_step'], d['train_step']['kernels_ms'])
PY
done
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu_31.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -3 gpurun_out/pytest_gpu_31.log
