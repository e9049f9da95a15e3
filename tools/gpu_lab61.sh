#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu_61.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -3 gpurun_out/pytest_gpu_61.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 500 python bench.py --workload volleyball_res18_lite128_T10_N12_720p --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_res18_61.json 2> gpurun_out/bench_res18_61.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_res18_61.json').read().strip().splitlines()[-1])
print(l['value']); print(json.dumps(l.get('train_step')))
PY
