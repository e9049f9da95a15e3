#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gpu.py tests/test_properties_gpu.py tests/test_e2e_gpu.py -m gpu -q 2>&1 | tail -3
for v in 0 1; do
DIN_CONV_RESIDENT=$v timeout 600 python bench.py --no-cpu-baseline --no-train-step --steps 10 > gpurun_out/bench_32_res$v.json 2> gpurun_out/bench_32.err
echo "bench resident=$v rc=$?"; python - <<PY
import json
d=json.load(open('gpurun_out/bench_32_res$v.json')); r=d['roofline']
print('clips/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],2), 'convTF', round(r['achieved']), 'frac', round(r['frac'],3), d['clocks'])
print({k:v for k,v in r['per_layer_tflops'].items() if '64->' in k or '128->128' in k})
PY
done
