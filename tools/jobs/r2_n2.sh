#!/bin/bash
# 2-GPU job: NCCL gradient exchange test, DataParallel on two devices, bench at N=2 (incl. the training step with its all-reduce)
O=gpurun_out/r2n2; mkdir -p $O
nvidia-smi -L
timeout 900 python -m pytest tests/test_ddp_gpu.py tests/test_dropin_gpu.py -m gpu -q --timeout 900 -p no:cacheprovider -rA -s > $O/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|skipped|world 2" $O/pytest.log | tail -5; grep -E "^FAILED|Error" $O/pytest.log | head
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"; head -c 300 $O/bench_n2.json; echo
grep -E "NVLS|Channel|via P2P|busbw|AllReduce" $O/bench_n2.err | head -20 > $O/nccl_trace_excerpt.txt
timeout 600 python bench.py --gpus 1 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench n1 rc=$?"
python - <<'PY'
import json
for f in ('n1','n2'):
    d=json.loads(open(f'gpurun_out/r2n2/bench_{f}.json').read().strip().splitlines()[-1])
    print(f, d['value'], d['e2e']['value'], d.get('train_step'))
PY
