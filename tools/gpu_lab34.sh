#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gpu.py -m gpu -q -k "packing" 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_34.json 2> gpurun_out/bench_34.err
echo "bench rc=$?"; python - <<PY
import json
d=json.load(open('gpurun_out/bench_34.json')); r=d['roofline']
print('clips/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],2), 'convTF', round(r['achieved']), 'frac', round(r['frac'],3), d['clocks'])
print('train_step', json.dumps(d['train_step']))
PY
tail -2 gpurun_out/bench_34.err
