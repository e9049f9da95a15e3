#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_bwd_gpu.py tests/test_backward_gpu.py -m gpu -q -x -k "fused_relu or full_training or optimizer or stage1 or basenet" > gpurun_out/pytest_55.log 2>&1
echo "rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest_55.log | cut -c1-250 | head -20
DIN_KINETO=0 timeout 300 python tests/tools/train_host_profile.py res18 > gpurun_out/host_prof_res18_d.log 2>&1; echo "rc=$?"
DIN_KINETO=0 timeout 300 python tests/tools/train_host_profile.py vgg16 > gpurun_out/host_prof_vgg16_d.log 2>&1; echo "rc=$?"
grep "host issue" gpurun_out/host_prof_*_d.log
