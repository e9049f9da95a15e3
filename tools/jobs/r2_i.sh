#!/bin/bash
# full GPU suite + every workload's bench line
O=gpurun_out/r2i; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -rA > $O/pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error" $O/pytest.log | tail -3
grep -E "^FAILED|^ERROR|degenerate|stem pair" $O/pytest.log | head -40
for w in volleyball_vgg16_lite128_T10_N12_720p volleyball_vgg16_hier_st_T10_N12_720p volleyball_res18_lite128_T10_N12_720p volleyball_inv3_full_T10_N12_720p collective_res18_T10_N13_480p; do
  timeout 600 python bench.py --workload $w > $O/bench_$w.json 2> $O/bench_$w.err; echo "bench $w rc=$?"; head -c 300 $O/bench_$w.json; echo
done
timeout 300 ncu --set full --clock-control none -k regex:"roi_align_kernel|dynamic_infer" -c 12 -o $O/ncu_head -f python tools/prof_head.py > $O/ncu_head.log 2>&1; echo "ncu rc=$?"
