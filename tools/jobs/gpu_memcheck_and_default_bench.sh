#!/bin/bash
# memcheck of the kernels added in the round's second half + default bench with the ingest key
O=gpurun_out/memcheck; mkdir -p $O
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_conv_bwd_gpu.py tests/test_conv_gpu.py tests/test_ingest_gpu.py -m gpu -q --timeout 1100 -p no:cacheprovider -k "general or inception_geometry or pad0 or relu_slice or branch_group or resize_kernel_is_bit or (wgrad_matches and not 720)" > $O/memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds" $O/memcheck.log | tail -8
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/memcheck/bench_default.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d.get('ingest'))
PY
