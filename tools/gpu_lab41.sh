#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_e2e_gpu.py tests/test_edge_cases_gpu.py tests/test_stage1_loss_gpu.py -m gpu -q 2>&1 | tail -3
for w in volleyball_inv3_full_T10_N12_720p volleyball_res18_lite128_T10_N12_720p; do
  timeout 900 python bench.py --steps 5 --warmup 3 --workload $w > gpurun_out/bench41_$w.json 2> gpurun_out/bench41.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench41_$w.json')); r=d['roofline']
print('$w', 'clips/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'e2e_u8', round(d['e2e_u8']['value'],1), 'ms', round(d['ms_per_step'],2), 'convTF', round(r['achieved']), 'whole_frac', round(r['whole_path_frac'],3), r['other_kernels_ms'])
PY
done
