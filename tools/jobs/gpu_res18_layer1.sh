#!/bin/bash
O=gpurun_out/l1; mkdir -p $O
python tools/probes/res18_layer1_probe.py
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_conv_bwd_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
for w in volleyball_inv3_full_T10_N12_720p volleyball_res18_lite128_T10_N12_720p volleyball_vgg16_lite128_T10_N12_720p; do
timeout 600 python bench.py --workload $w --no-cpu-baseline --no-train-step --no-ingest > $O/b_$w.json 2> $O/b_$w.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/l1/b_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
    print(f, round(d['value'],1), 'e2e', round(d['e2e']['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], {k:round(v,2) for k,v in list(r['per_layer_ms_per_step'].items())[:4]})
PY
