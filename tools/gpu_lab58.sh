#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu_58.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -3 gpurun_out/pytest_gpu_58.log
timeout 600 python bench.py > gpurun_out/bench_default_58.json 2> gpurun_out/bench_default_58.err; echo "bench rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_default_58.json').read().strip().splitlines()[-1])
print({k: l[k] for k in ('metric','value','ms_per_step','gpu_launches')}, l['e2e']['value'], (l.get('e2e_u8') or {}).get('value'))
print(l['roofline']); print(l['cpu_baseline']); print(l['clocks'])
print(json.dumps(l.get('train_step')))
PY
timeout 700 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_conv_bwd_gpu.py tests/test_conv_gpu.py -m gpu -q -x -k "maxpool3s2 or scatter2 or bn_gamma or batch_stat or fused_relu or batched_packing or stride2" > gpurun_out/memcheck_58.log 2>&1
echo "memcheck rc=$?"; grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds" gpurun_out/memcheck_58.log | head -8
