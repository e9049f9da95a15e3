#!/bin/bash
O=gpurun_out/r2o; mkdir -p $O
timeout 900 python -m pytest tests/test_e2e_gpu.py tests/test_fullsize_gpu.py -m gpu -q -k "tce" --timeout 600 -p no:cacheprovider -s > $O/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|\[e2e\]|Error" $O/pytest.log | tail -8
timeout 600 python bench.py --workload volleyball_vgg16_tce_T10_N12_720p > $O/bench_tce.json 2> $O/bench_tce.err; echo "bench rc=$?"; tail -3 $O/bench_tce.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2o/bench_tce.json').read().strip().splitlines()[-1]); r=d['roofline']
print(d['value'], d['e2e']['value'], d['ms_per_step'], r['other_kernels_ms'], d['cpu_baseline'])
PY
