#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -f -o gpurun_out/prof_conv_v3 \
   python tools/prof_conv.py > gpurun_out/ncu_v3.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_v3.log; ls -la gpurun_out/*.ncu-rep
