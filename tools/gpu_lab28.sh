#!/bin/bash
mkdir -p gpurun_out
for w in volleyball_inv3_full_T10_N12_720p volleyball_res18_lite128_T10_N12_720p; do
  timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload $w > gpurun_out/bench28_$w.json 2> gpurun_out/bench28_$w.err
  echo "$w rc=$?"; tail -2 gpurun_out/bench28_$w.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench28_$w.json')); r=d['roofline']
print('$w', 'clips/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],2), 'convTF', round(r['achieved']), 'conv_ms', round(r['kernel_ms_per_step'],2), 'all_ms', round(r['all_kernels_ms_per_step'],2), r['other_kernels_ms'], 'launches', d['gpu_launches'], 'whole_frac', round(r['whole_path_frac'],3))
pl=sorted(r['per_layer_tflops'].items(), key=lambda kv: kv[1])[:8]
print('slowest layers', pl)
PY
done
