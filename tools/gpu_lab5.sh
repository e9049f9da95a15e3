#!/bin/bash
# full GPU test suite + bench (no ncu)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" ; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench_cur.json 2> gpurun_out/bench_cur.err
echo "bench rc=$?"; cat gpurun_out/bench_cur.json; tail -5 gpurun_out/bench_cur.err
