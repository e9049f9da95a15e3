#!/bin/bash
# ncu --set full of the ResNet-18 training-step kernels added in v12 (one step, BatchNorm eval mode and batch statistics)
mkdir -p gpurun_out
DIN_NCU=1 timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:"maxpool3s2|bn_|scatter2|add_f16|pack_weight|scale_rows|stem_wgrad|stem_tc_wide" -s 30 -c 45 -f -o gpurun_out/prof_train_res18_v12 \
  python tests/tools/train_host_profile.py res18 > gpurun_out/ncu60a.log 2>&1; echo "rc=$?"
DIN_NCU=1 timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:"bn_stats|bn_finalize|bn_apply|bn_bwd" -s 42 -c 40 -f -o gpurun_out/prof_train_res18_bn_v12 \
  python tests/tools/train_host_profile.py res18 bn > gpurun_out/ncu60b.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py gpurun_out/prof_train_res18_v12.ncu-rep gpurun_out/ncu_r1_v12_res18_train_kernels.md --json gpurun_out/ncu_r1_v12_res18_train_kernels.json
python tools/ncu_summary.py gpurun_out/prof_train_res18_bn_v12.ncu-rep gpurun_out/ncu_r1_v12_res18_bn_kernels.md --json gpurun_out/ncu_r1_v12_res18_bn_kernels.json
ls -la gpurun_out/*.ncu-rep | tail -3; head -12 gpurun_out/ncu_r1_v12_res18_train_kernels.md | cut -c1-220
