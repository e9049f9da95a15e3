#!/bin/bash
mkdir -p gpurun_out
timeout 500 python bench.py --workload volleyball_res18_lite128_T10_N12_720p --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_res18_train.json 2> gpurun_out/bench_res18_train.err
echo "rc=$?"; python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_res18_train.json').read().strip().splitlines()[-1])
print(l['value'], l.get('e2e'))
print(json.dumps(l.get('train_step'), indent=1))
PY
tail -3 gpurun_out/bench_res18_train.err
