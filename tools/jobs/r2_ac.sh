#!/bin/bash
# host-memory frame streaming: test + e2e in the default bench and two more workloads
O=gpurun_out/r2ac; mkdir -p $O
timeout 600 python -m pytest tests/test_ingest_gpu.py tests/test_edge_cases_gpu.py tests/test_dropin_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -rA > $O/pytest.log 2>&1; echo "rc=$?"
grep -E "passed|failed|error" $O/pytest.log | tail -3; grep -E "^FAILED|^ERROR|^E  " $O/pytest.log | head -20
for w in volleyball_vgg16_lite128_T10_N12_720p volleyball_res18_lite128_T10_N12_720p collective_res18_T10_N13_480p; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --no-train-step --no-ingest > $O/bench_$w.json 2> $O/bench_$w.err; echo "bench $w rc=$?"; tail -2 $O/bench_$w.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2ac/bench_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(d['config']['workload'], round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'u8', round(d['e2e_u8']['value'],1), d['clocks']['sm_mhz'])
PY
