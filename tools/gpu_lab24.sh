#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tests/tools/train_ddp_check.py > gpurun_out/train_ddp_n2.log 2>&1
echo "ddp rc=$?"; grep -v Warning gpurun_out/train_ddp_n2.log | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_24.json 2> gpurun_out/bench_n2_24.err
echo "n2 rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n2_24.json'))
print('n_gpus', d['n_gpus'], 'clips/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],2), d['clocks'], 'launches', d['gpu_launches'])
PY
tail -2 gpurun_out/bench_n2_24.err
