#!/bin/bash
# evidence for rows a3 / a4 (ResNet-18, Inception-v3): bench lines (not under a profiler) + ncu launch lists
mkdir -p gpurun_out
for w in volleyball_inv3_full_T10_N12_720p volleyball_res18_lite128_T10_N12_720p; do
  timeout 900 python bench.py --steps 5 --warmup 3 --workload $w > gpurun_out/bench38_$w.json 2> gpurun_out/bench38_$w.err
  echo "$w rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/bench38_$w.json')); r=d['roofline']
print('$w', 'clips/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'e2e_u8', round(d['e2e_u8']['value'],1), 'ms', round(d['ms_per_step'],2), 'convTF', round(r['achieved']), 'whole_frac', round(r['whole_path_frac'],3), r['other_kernels_ms'], 'cpu', d['cpu_baseline']['value'])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r1_v11_inv3.csv \
   python bench.py --steps 1 --warmup 3 --clips-per-gpu 2 --no-cpu-baseline --no-e2e --workload volleyball_inv3_full_T10_N12_720p > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r1_v11_res18.csv \
   python bench.py --steps 1 --warmup 3 --clips-per-gpu 2 --no-cpu-baseline --no-e2e --workload volleyball_res18_lite128_T10_N12_720p > /dev/null 2>&1
wc -l gpurun_out/launches_r1_v11_*.csv
