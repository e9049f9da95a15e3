#!/bin/bash
# CTA-pair wgrad: parity + training-step A/B
O=gpurun_out/r2z; mkdir -p $O
timeout 600 python -m pytest tests/test_conv_bwd_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -rA -s -k "wgrad" > $O/pytest.log 2>&1; echo "rc=$?"
grep -E "passed|failed|error" $O/pytest.log | tail -3; grep -E "^FAILED|^ERROR|^E  " $O/pytest.log | head -20
for m in 1 0; do
  DIN_WGRAD_2CTA=$m timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 3 > $O/bench_wg$m.json 2> $O/bench_wg$m.err; echo "bench 2cta=$m rc=$?"
done
python - <<'PY'
import json
for m in (1,0):
    d=json.loads(open(f'gpurun_out/r2z/bench_wg{m}.json').read().strip().splitlines()[-1])
    t=d['train_step']; print(m, t.get('ms_per_step'), t.get('kernels_ms'), t.get('algorithmic_tflops_per_gpu'), t.get('error'))
PY
