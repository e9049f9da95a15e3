#!/bin/bash
# launch list of the default bench command + ncu --set full of one forward step's convolution kernels (20-frame chunk)
O=gpurun_out/ncuf; mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_r2_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train-step --no-ingest --no-e2e > $O/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_igemm|conv1_fused" --launch-skip 36 -c 12 -o $O/ncu_r2_final_step -f python bench.py --clips-per-gpu 2 --steps 1 --warmup 3 --no-cpu-baseline --no-train-step --no-ingest --no-e2e > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 $O/ncu_full.log
python tools/ncu_summary.py $O/ncu_r2_final_step.ncu-rep $O/ncu_r2_final_step.md --json $O/ncu_r2_final_step.json; cat $O/ncu_r2_final_step.md
