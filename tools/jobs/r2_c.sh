#!/bin/bash
# round-2 GPU job C: the fused conv1_1 + conv1_2 kernel -- correctness first (bounded by timeout), then bench A/B
O=gpurun_out/r2c; mkdir -p $O
timeout 600 python -m pytest tests/test_conv_gpu.py -m gpu -q -k "stem_pair" --timeout 300 -p no:cacheprovider -rA -s > $O/pytest_pair.log 2>&1; rc=$?; echo "pair tests rc=$rc"
grep -E "stem pair|passed|failed|Error|error" $O/pytest_pair.log | head -30
if [ $rc -ne 0 ]; then exit 0; fi
timeout 900 python -m pytest tests/test_e2e_gpu.py tests/test_fullsize_gpu.py tests/test_ingest_gpu.py tests/test_backward_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > $O/pytest_e2e.log 2>&1; echo "e2e rc=$?"
grep -E "passed|failed" $O/pytest_e2e.log | tail -2; grep -E "^FAILED|^\[e2e\]" $O/pytest_e2e.log | head -40
timeout 600 python bench.py --no-cpu-baseline > $O/bench_fused.json 2> $O/bench_fused.err; echo "bench fused rc=$?"; head -c 400 $O/bench_fused.json; echo
DIN_FUSE_CONV1=0 timeout 600 python bench.py --no-cpu-baseline --no-train-step > $O/bench_unfused.json 2> $O/bench_unfused.err; echo "bench unfused rc=$?"; head -c 400 $O/bench_unfused.json; echo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv1_fused" -c 2 -o $O/ncu_fused -f python bench.py --no-cpu-baseline --no-train-step --no-e2e --steps 1 --warmup 3 > $O/ncu_fused.log 2>&1; echo "ncu rc=$?"
