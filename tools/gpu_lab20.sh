#!/bin/bash
# ncu evidence for the round-1 session-2 kernels: full-set capture of the training kernels + a launch list of one
# bench run (forward steps + the supplementary training step)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_wgrad|relu_pool|stem_wgrad|stem_tc|din_bwd|din_conv_bwd|dynamic_infer' \
   -o gpurun_out/ncu_r1_v9_train python tools/prof_train.py > gpurun_out/ncu_train.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_train.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r1_v9.csv \
   python bench.py --steps 1 --warmup 3 --clips-per-gpu 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_r1_v9.csv
ls -la gpurun_out/*.ncu-rep
