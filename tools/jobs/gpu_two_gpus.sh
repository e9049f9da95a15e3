#!/bin/bash
# final tree on 2 GPUs: NCCL gradient exchange test, DataParallel, bench at N = 2 (forward, e2e with host streaming, train step with all-reduce)
O=gpurun_out/n2; mkdir -p $O
nvidia-smi -L
timeout 900 python -m pytest tests/test_ddp_gpu.py tests/test_dropin_gpu.py -m gpu -q --timeout 900 -p no:cacheprovider -rA > $O/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|skipped" $O/pytest.log | tail -3; grep -E "^FAILED|Error" $O/pytest.log | head
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"; tail -3 $O/bench_n2.err
timeout 600 python bench.py --gpus 1 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench n1 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err; echo "ref n2 rc=$?"; head -c 300 $O/bench_ref_n2.json
python - <<'PY'
import json
for f in ('n1','n2'):
    d=json.loads(open(f'gpurun_out/n2/bench_{f}.json').read().strip().splitlines()[-1])
    t=d.get('train_step') or {}
    print(f, round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'train', t.get('ms_per_step'), t.get('allreduce'), t.get('error'))
PY
