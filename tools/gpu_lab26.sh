#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_bwd_gpu.py tests/test_backward_gpu.py tests/test_stage1_loss_gpu.py -m gpu -q -k "relu_pool or stem_wgrad or full_training or stage1_training or per_clip" 2>&1 | tail -12 > gpurun_out/pytest_26.log
echo "rc=${PIPESTATUS[0]}"; tail -6 gpurun_out/pytest_26.log
for fpc in 16 80; do
DIN_FRAMES_PER_CHUNK=$fpc timeout 600 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_26_$fpc.json 2> gpurun_out/bench_26.err
echo "bench fpc=$fpc rc=$?"; python - <<PY
import json
d=json.load(open('gpurun_out/bench_26_$fpc.json')); r=d['roofline']
print('clips/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],2), 'convTF', round(r['achieved']), r['other_kernels_ms'], d['clocks'])
print('train_step', json.dumps(d.get('train_step')))
PY
done
tail -3 gpurun_out/bench_26.err
