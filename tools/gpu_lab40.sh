#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s -k "inception or inv3" 2>&1 | grep -E "passed|failed|\[e2e\]|\[basenet\]|^E  " | cut -c1-200
timeout 900 python bench.py --steps 5 --warmup 3 --workload volleyball_inv3_full_T10_N12_720p > gpurun_out/bench40_inv3.json 2> gpurun_out/bench40.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench40_inv3.json')); r=d['roofline']
print('inv3 clips/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'e2e_u8', round(d['e2e_u8']['value'],1), 'ms', round(d['ms_per_step'],2), 'convTF', round(r['achieved']), 'whole_frac', round(r['whole_path_frac'],3), r['other_kernels_ms'])
PY
