#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_ingest_gpu.py tests/test_e2e_gpu.py -m gpu -q -k "stem or res18 or u8" 2>&1 | tail -2
w=volleyball_res18_lite128_T10_N12_720p
for v in 0 1; do
DIN_STEM_WIDE=$v timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload $w > gpurun_out/bench43_$v.json 2> gpurun_out/bench43.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench43_$v.json')); r=d['roofline']
print('wide=$v', 'clips/s', round(d['value'],1), 'e2e_u8', round(d['e2e_u8']['value'],1), 'ms', round(d['ms_per_step'],2), r['other_kernels_ms'], d['clocks']['sm_mhz'])
PY
done
