#!/bin/bash
mkdir -p gpurun_out
timeout 45 python bench.py --workload volleyball_res18_lite128_T10_N12_720p --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --clips-per-gpu 8 > gpurun_out/bench_res18_67.json 2> gpurun_out/bench_res18_67.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_res18_67.json').read().strip().splitlines()[-1])
print(l['value']); print(json.dumps(l.get('train_step')))
PY
