#!/bin/bash
O=gpurun_out/r2l; mkdir -p $O
for v in ws 0; do
  if [ $v = ws ]; then unset DIN_STEM_S2D; else export DIN_STEM_S2D=0; fi
  timeout 600 python bench.py --workload volleyball_res18_lite128_T10_N12_720p --no-cpu-baseline --no-e2e > $O/bench_res18_$v.json 2> $O/bench_res18_$v.err
done
unset DIN_STEM_S2D
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2l/bench_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); t=d['train_step']
    print(f.split('/')[-1], d['value'], t['ms_per_step'], t['kernels_ms'])
PY
timeout 600 ncu --set full --clock-control none -k regex:"stem_s2d_ws" -c 1 -o $O/ncu_stem -f python bench.py --workload volleyball_res18_lite128_T10_N12_720p --no-cpu-baseline --no-train-step --no-e2e --steps 1 --warmup 3 > $O/ncu_stem.log 2>&1; echo "ncu rc=$?"
