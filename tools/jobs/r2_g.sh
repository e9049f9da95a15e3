#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
timeout 300 python -m pytest tests/test_conv_gpu.py -m gpu -q -k stem_pair --timeout 300 -p no:cacheprovider > $O/pytest_a.log 2>&1; rc=$?; echo "pair tests rc=$rc"; grep -E "passed|failed" $O/pytest_a.log | tail -1
timeout 600 python bench.py --no-cpu-baseline --no-train-step --no-e2e > $O/bench_fused.json 2> $O/bench_fused.err; echo "bench fused rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2g/bench_fused.json').read().strip().splitlines()[-1]); r=d['roofline']
print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz']); print({k:v for k,v in r['per_layer_tflops'].items() if '->64' in k or '64->' in k})
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv1_fused" -c 1 -o $O/ncu_fused -f python bench.py --no-cpu-baseline --no-train-step --no-e2e --steps 1 --warmup 3 > $O/ncu_fused.log 2>&1; echo "ncu rc=$?"
