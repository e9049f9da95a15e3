#!/bin/bash
O=gpurun_out/r2pair; mkdir -p $O
timeout 600 python -m pytest tests/test_conv_bwd_gpu.py tests/test_backward_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -k "wgrad or full_training_step" > $O/pytest.log 2>&1; echo "rc=$?"; grep -E "passed|failed" $O/pytest.log | tail -2; grep -E "^FAILED|^E  " $O/pytest.log | head
for m in 1 0; do
  DIN_WGRAD_PAIR64=$m timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-ingest --steps 3 > $O/b$m.json 2> $O/b$m.err
  python - <<PY
import json
d=json.loads(open('$O/b$m.json').read().strip().splitlines()[-1]); t=d['train_step']
print('pair64=$m', t['ms_per_step'], t['kernels_ms'])
PY
done
