#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_igemm|stem_tc" -f -o gpurun_out/prof_v4 \
   python tools/prof_conv.py > gpurun_out/ncu_v4.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_v4.log; ls -la gpurun_out/*.ncu-rep
