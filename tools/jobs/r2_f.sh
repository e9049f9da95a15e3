#!/bin/bash
O=gpurun_out/r2f; mkdir -p $O
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_head_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider > $O/pytest_a.log 2>&1; rc=$?; echo "conv+head tests rc=$rc"
grep -E "passed|failed" $O/pytest_a.log | tail -2; grep -E "^FAILED|Error" $O/pytest_a.log | head
if [ $rc -ne 0 ]; then DIN_FUSED_DEBUG=1 timeout 120 python tools/debug_fused.py 2 48 64; fi
timeout 600 python bench.py --no-cpu-baseline --no-train-step --no-e2e > $O/bench_fused.json 2> $O/bench_fused.err; echo "bench fused rc=$?"
DIN_FUSE_CONV1=0 timeout 600 python bench.py --no-cpu-baseline --no-train-step --no-e2e > $O/bench_unfused.json 2> $O/bench_unfused.err; echo "bench unfused rc=$?"
python - <<'PY'
import json
for f in ('fused','unfused'):
    try:
        d=json.loads(open(f'gpurun_out/r2f/bench_{f}.json').read().strip().splitlines()[-1]); r=d['roofline']
        print(f, d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], r['other_kernels_ms']); print({k:v for k,v in r['per_layer_tflops'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
timeout 300 ncu --set full --clock-control none -k regex:"roi_align_kernel|dynamic_infer" -c 12 -o $O/ncu_head -f python tools/prof_head.py > $O/ncu_head.log 2>&1; echo "ncu rc=$?"; tail -2 $O/ncu_head.log
