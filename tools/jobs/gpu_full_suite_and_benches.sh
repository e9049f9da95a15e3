#!/bin/bash
# full GPU suite + smoke + benches of every workload (profiles/bench_r2_c_*.json)
O=gpurun_out/full; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -rA > $O/pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error" $O/pytest.log | tail -3
grep -E "^FAILED|^ERROR" $O/pytest.log | head -40
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
for w in volleyball_vgg16_lite128_T10_N12_720p volleyball_inv3_full_T10_N12_720p volleyball_res18_lite128_T10_N12_720p collective_res18_T10_N13_480p volleyball_vgg16_hier_st_T10_N12_720p volleyball_vgg16_tce_T10_N12_720p; do
  timeout 600 python bench.py --workload $w > $O/bench_$w.json 2> $O/bench_$w.err; echo "bench $w rc=$?"; head -c 160 $O/bench_$w.json; echo
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/full/bench_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
    print(d['config']['workload'], round(d['value'],1), round(d['e2e']['value'],1), round(d.get('e2e_u8',{}).get('value',0),1), round(d['ms_per_step'],2), round(r['frac'],3), round(r['whole_path_frac'],3), d['clocks']['sm_mhz'], (d.get('train_step') or {}).get('ms_per_step'), (d.get('train_step') or {}).get('error'))
PY
