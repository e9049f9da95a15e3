#!/bin/bash
# merged Inception branch heads: kernel tests, plan A/B, e2e parity, bench A/B
O=gpurun_out/r2s; mkdir -p $O
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_e2e_gpu.py tests/test_fullsize_gpu.py tests/test_backward_gpu.py -m gpu -q --timeout 900 -p no:cacheprovider -rA -k "branch or inception or inv3 or stem_and_pool or batchnorm_on_batch or fixture" > $O/pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error" $O/pytest.log | tail -3
grep -E "^FAILED|^ERROR|\[e2e\]" $O/pytest.log | head -40
for m in 1 0; do
  DIN_INV3_MERGE=$m timeout 600 python bench.py --workload volleyball_inv3_full_T10_N12_720p --no-cpu-baseline > $O/bench_inv3_merge$m.json 2> $O/bench_inv3_merge$m.err; echo "bench merge=$m rc=$?"
done
python - <<'PY'
import json
for m in (1,0):
    d=json.loads(open(f'gpurun_out/r2s/bench_inv3_merge{m}.json').read().strip().splitlines()[-1]); r=d['roofline']
    print(m, d['value'], d['e2e']['value'], d['ms_per_step'], r['kernel_ms_per_step'], r['other_kernels_ms'])
    print({k:v for k,v in r['per_layer_tflops'].items() if '1x1' in k})
PY
