#!/bin/bash
O=gpurun_out/r2j; mkdir -p $O
timeout 600 python -m pytest tests/test_head_gpu.py tests/test_e2e_gpu.py tests/test_backward_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" $O/pytest.log | tail -2; grep -E "^FAILED" $O/pytest.log | head
timeout 300 ncu --set full --clock-control none -k regex:"roi_align_kernel|dynamic_infer" -c 12 -o $O/ncu_head -f python tools/prof_head.py > $O/ncu_head.log 2>&1; echo "ncu rc=$?"
timeout 600 python bench.py --workload volleyball_res18_lite128_T10_N12_720p --no-cpu-baseline --no-e2e > $O/bench_res18.json 2> $O/bench_res18.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2j/bench_res18.json').read().strip().splitlines()[-1])
print(d['value'], d['train_step']['ms_per_step'])
PY
