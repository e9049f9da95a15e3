#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_ingest_gpu.py tests/test_e2e_gpu.py tests/test_edge_cases_gpu.py -m gpu -q -k "stem or u8 or e2e or small or config1 or inception or collective or edge or degenerate or wide or odd" 2>&1 | tail -8 > gpurun_out/pytest_21.log
echo "rc=${PIPESTATUS[0]}"; tail -4 gpurun_out/pytest_21.log
for v in 0 1; do
DIN_STEM_PIPE=$v timeout 600 python bench.py --no-cpu-baseline --no-train-step --steps 10 > gpurun_out/bench_21_pipe$v.json 2> gpurun_out/bench_21.err
echo "bench pipe=$v rc=$?"; python - <<PY
import json
d=json.load(open('gpurun_out/bench_21_pipe$v.json')); r=d['roofline']
print('clips/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],2), 'convTF', round(r['achieved']), r['other_kernels_ms'], d['clocks'])
PY
done
tail -3 gpurun_out/bench_21.err
