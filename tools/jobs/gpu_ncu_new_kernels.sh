#!/bin/bash
O=gpurun_out/ncu; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"wgrad|conv_igemm|avgpool3s1|upsample|maxpool3s2|resample" -o $O/ncu_r2b -f python tools/prof_r2b.py > $O/ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $O/ncu.log
python tools/ncu_summary.py $O/ncu_r2b.ncu-rep $O/ncu_r2b.md --json $O/ncu_r2b.json; cat $O/ncu_r2b.md
ls -la $O
