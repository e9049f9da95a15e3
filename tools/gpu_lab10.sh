#!/bin/bash
mkdir -p gpurun_out
for c in 16 40 80; do
  DIN_FRAMES_PER_CHUNK=$c timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_chunk$c.json 2>gpurun_out/bench_chunk$c.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_chunk$c.json')); print('chunk', $c, 'clips/s', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'convTF', round(d['roofline']['achieved']), d['roofline']['other_kernels_ms'].get('stem3x3s1'), d['clocks']['sm_mhz'])
PY
done
