#!/bin/bash
mkdir -p gpurun_out
DIN_KINETO=0 timeout 300 python tests/tools/train_host_profile.py res18 bn > gpurun_out/host_prof_res18_bn.log 2>&1; echo "rc=$?"
grep "host issue" gpurun_out/host_prof_res18_bn.log; tail -3 gpurun_out/host_prof_res18_bn.log | cut -c1-200
timeout 500 python bench.py --workload volleyball_res18_lite128_T10_N12_720p --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_res18_train.json 2> gpurun_out/bench_res18_train.err
echo "rc=$?"; python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_res18_train.json').read().strip().splitlines()[-1])
print(l['value'], l.get('e2e', {}).get('value'), (l.get('e2e_u8') or {}).get('value'))
print(json.dumps(l.get('train_step')))
PY
