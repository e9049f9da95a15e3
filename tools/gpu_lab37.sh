#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 > gpurun_out/bench_37.json 2> gpurun_out/bench_37.err
echo "bench rc=$?"; python - <<PY
import json
d=json.load(open('gpurun_out/bench_37.json')); r=d['roofline']
print('clips/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'e2e_u8', round(d['e2e_u8']['value'],1), d['e2e_u8']['h2d_bytes_per_step'], 'ms', round(d['ms_per_step'],2), 'convTF', round(r['achieved']), 'frac', round(r['frac'],3), d['clocks'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
PY
tail -2 gpurun_out/bench_37.err
