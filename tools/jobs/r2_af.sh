#!/bin/bash
# host-frame staging slots probe, window relu/pool backward test + train-step A/B, host profile of the Inception-v3 training step
O=gpurun_out/r2af; mkdir -p $O
for n in 3 6; do echo SLOTS $n; DIN_STAGE_SLOTS=$n python tools/probes/e2e_probe.py 2>&1 | grep -E "^host|^device"; done
python -m pytest tests/test_conv_bwd_gpu.py -m gpu -q -k relu_pool -p no:cacheprovider 2>&1 | tail -2
for m in 0 1; do DIN_RELU_POOL_PIXEL=$m python bench.py --no-cpu-baseline --no-e2e --no-ingest --steps 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); t=d['train_step']; print('pixel=$m', t['ms_per_step'], t['kernels_ms'].get('relu'))"; done
python tools/probes/train_profile.py > $O/train_profile_inv3.txt 2>&1; head -70 $O/train_profile_inv3.txt
