#!/bin/bash
# stage-1 Inception-v3 training, inv3 bench incl. train step; fc_emb split A/B
O=gpurun_out/r2v; mkdir -p $O
timeout 900 python -m pytest tests/test_stage1_loss_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -rA -s -k "resnet18" > $O/pytest.log 2>&1; echo "stage1 rc=$?"
grep -E "passed|failed|error" $O/pytest.log | tail -3; grep -E "^FAILED|^ERROR|^E  |\[stage1" $O/pytest.log | head -30
timeout 600 python bench.py --workload volleyball_inv3_full_T10_N12_720p --no-cpu-baseline > $O/bench_inv3.json 2> $O/bench_inv3.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2v/bench_inv3.json').read().strip().splitlines()[-1]); r=d['roofline']
print(d['value'], d['e2e']['value'], d['ms_per_step'], r['kernel_ms_per_step'], r['other_kernels_ms'])
print(d.get('train_step'))
PY
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
