#!/bin/bash
O=gpurun_out/r2e; mkdir -p $O
timeout 600 python -m pytest tests/test_conv_gpu.py -m gpu -q -k "stem_pair" --timeout 300 -p no:cacheprovider -rA -s > $O/pytest_pair.log 2>&1; rc=$?; echo "pair tests rc=$rc"
grep -E "stem pair|passed|failed|Error|error" $O/pytest_pair.log | head -12
if [ $rc -ne 0 ]; then DIN_FUSED_DEBUG=1 timeout 120 python tools/debug_fused.py 2 48 64; exit 0; fi
timeout 600 python bench.py --no-cpu-baseline --no-train-step --no-e2e > $O/bench_fused.json 2> $O/bench_fused.err; echo "bench fused rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2e/bench_fused.json').read().strip().splitlines()[-1]); r=d['roofline']
print(d['value'], d['ms_per_step'], d['clocks']); print({k:v for k,v in r['per_layer_tflops'].items() if '3->64' in k or '64->128' in k})
PY
