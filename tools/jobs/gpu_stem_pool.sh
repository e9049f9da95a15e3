#!/bin/bash
# fused ResNet-18 stem + max-pool: probe, bit-identity tests, ResNet-18 parity (e2e / full size / u8), benches
O=gpurun_out/stempool; mkdir -p $O
python tools/probes/stem_probe.py
timeout 1200 python -m pytest tests/test_conv_gpu.py tests/test_e2e_gpu.py tests/test_fullsize_gpu.py tests/test_ingest_gpu.py tests/test_stage1_loss_gpu.py -m gpu -q --timeout 900 -p no:cacheprovider -rA -k "stem or fused_maxpool or res18 or collective or u8 or basenet" > $O/pytest.log 2>&1; echo "rc=$?"
grep -E "passed|failed|error" $O/pytest.log | tail -3; grep -E "^FAILED|^ERROR|^E  " $O/pytest.log | head -30
for w in volleyball_res18_lite128_T10_N12_720p collective_res18_T10_N13_480p; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --no-train-step --no-ingest > $O/b_${w}.json 2> $O/b_${w}.err
  python - <<PY
import json
d=json.loads(open('$O/b_${w}.json').read().strip().splitlines()[-1]); r=d['roofline']
print('$w', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'u8', round(d['e2e_u8']['value'],1), round(d['ms_per_step'],2), r['other_kernels_ms'], d['clocks']['sm_mhz'])
PY
done
