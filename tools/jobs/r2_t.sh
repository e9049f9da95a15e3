#!/bin/bash
# k16_last (partial last K block), BN cost model, strip upsample, fc_emb single-part weights: conv tests + e2e + fullsize parity, bench inv3 / headline / res18
O=gpurun_out/r2t; mkdir -p $O
timeout 1200 python -m pytest tests/test_conv_gpu.py tests/test_e2e_gpu.py tests/test_fullsize_gpu.py tests/test_edge_cases_gpu.py -m gpu -q --timeout 900 -p no:cacheprovider -rA > $O/pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error" $O/pytest.log | tail -3
grep -E "^FAILED|^ERROR|\[e2e\]" $O/pytest.log | head -40
for w in volleyball_inv3_full_T10_N12_720p volleyball_vgg16_lite128_T10_N12_720p volleyball_res18_lite128_T10_N12_720p; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --no-train-step > $O/bench_$w.json 2> $O/bench_$w.err; echo "bench $w rc=$?"
done
python - <<'PY'
import json
for w in ['volleyball_inv3_full_T10_N12_720p','volleyball_vgg16_lite128_T10_N12_720p','volleyball_res18_lite128_T10_N12_720p']:
    d=json.loads(open(f'gpurun_out/r2t/bench_{w}.json').read().strip().splitlines()[-1]); r=d['roofline']
    print(w, d['value'], d['e2e']['value'], d['ms_per_step'], r['kernel_ms_per_step'], r['other_kernels_ms'], d['clocks'])
    print(sorted(r['per_layer_ms_per_step'].items(), key=lambda kv:-kv[1])[:14])
PY
