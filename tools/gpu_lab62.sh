#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_v12.json 2> gpurun_out/bench_n2_v12.err
echo "n2 rc=$?"; python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_n2_v12.json').read().strip().splitlines()[-1])
print({k: l[k] for k in ('value','n_gpus','ms_per_step','scaling')}, l['e2e']['value'], l['clocks'])
PY
tail -2 gpurun_out/bench_n2_v12.err
