#!/bin/bash
# session-3 baseline: full GPU suite + smoke + default bench + inv3 bench
O=gpurun_out/r2r; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -rA > $O/pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error" $O/pytest.log | tail -3
grep -E "^FAILED|^ERROR" $O/pytest.log | head -40
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
for w in volleyball_vgg16_lite128_T10_N12_720p volleyball_inv3_full_T10_N12_720p; do
  timeout 600 python bench.py --workload $w > $O/bench_$w.json 2> $O/bench_$w.err; echo "bench $w rc=$?"; head -c 200 $O/bench_$w.json; echo
done
