#!/bin/bash
O=gpurun_out/r2m; mkdir -p $O
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_e2e_gpu.py tests/test_ingest_gpu.py tests/test_fullsize_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" $O/pytest.log | tail -2; grep -E "^FAILED" $O/pytest.log | head
for w in volleyball_res18_lite128_T10_N12_720p collective_res18_T10_N13_480p; do
timeout 600 python bench.py --workload $w --no-cpu-baseline --no-train-step > $O/bench_$w.json 2> $O/bench_$w.err; echo "bench rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2m/bench_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
    print(f.split('/')[-1], d['value'], d['e2e']['value'], d['e2e_u8']['value'], d['ms_per_step'], d['clocks']['sm_mhz'], r['other_kernels_ms'])
PY
