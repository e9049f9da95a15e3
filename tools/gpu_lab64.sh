#!/bin/bash
# small ncu --set full capture (report deleted after summarising: gpurun_out copy-back limit is 64 MiB)
mkdir -p gpurun_out
DIN_NCU=1 timeout 170 ncu --set full --clock-control none --import-source on \
  -k regex:"pack_weight_kernel|maxpool3s2_relu_bwd|bn_stats_kernel|bn_apply_kernel" -s 2 -c 6 -f -o gpurun_out/prof_v12_small \
  python tests/tools/train_host_profile.py res18 bn > gpurun_out/ncu64.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py gpurun_out/prof_v12_small.ncu-rep gpurun_out/ncu_r1_v12_pack_pool_bn_full.md --json gpurun_out/ncu_r1_v12_pack_pool_bn_full.json
rm -f gpurun_out/*.ncu-rep
cat gpurun_out/ncu_r1_v12_pack_pool_bn_full.md | cut -c1-230; tail -3 gpurun_out/ncu64.log | cut -c1-200
