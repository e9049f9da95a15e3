#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_ingest_gpu.py tests/test_conv_bwd_gpu.py tests/test_backward_gpu.py tests/test_e2e_gpu.py -m gpu -q -k "stem or u8 or relu_pool or dynamic_infer_bwd or small or training_step_matches_oracle or inception or full_training" 2>&1 | tail -8 > gpurun_out/pytest_22.log
echo "rc=${PIPESTATUS[0]}"; tail -4 gpurun_out/pytest_22.log
timeout 600 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_22.json 2> gpurun_out/bench_22.err
echo "bench rc=$?"; python - <<PY
import json
d=json.load(open('gpurun_out/bench_22.json')); r=d['roofline']
print('clips/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],2), 'convTF', round(r['achieved']), r['other_kernels_ms'], d['clocks'])
print('train_step', json.dumps(d.get('train_step')))
PY
tail -3 gpurun_out/bench_22.err
