#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_igemm|stem_tc" -f -o gpurun_out/prof_v6 \
   python tools/prof_conv.py > gpurun_out/ncu_v6.log 2>&1
echo "ncu full rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r1_v6.csv \
   python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --clips-per-gpu 2 > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"; ls -la gpurun_out/ | tail -5
