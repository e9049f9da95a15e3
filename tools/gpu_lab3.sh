#!/bin/bash
# GPU lab run 3: full pytest -m gpu in one process, first bench line, ncu launch list + full capture.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1a.json 2> gpurun_out/bench_r1a.err
echo "bench rc=$?"; cat gpurun_out/bench_r1a.json; tail -5 gpurun_out/bench_r1a.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1a.csv \
   python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --clips-per-gpu 2 > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 24 -c 12 -f -o gpurun_out/prof_conv_r1a \
   python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --clips-per-gpu 2 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/
