#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
echo "n8 rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n8.json'))
print('n_gpus', d['n_gpus'], 'clips/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],2), d['clocks'], 'launches', d['gpu_launches'])
PY
tail -2 gpurun_out/bench_n8.err | cut -c1-300
