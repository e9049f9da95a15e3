#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_backward_gpu.py tests/test_e2e_gpu.py -m gpu -q -x -k "pack or full_training or optimizer or lite" > gpurun_out/pytest_54.log 2>&1
echo "rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest_54.log | cut -c1-250 | head -20
DIN_KINETO=0 timeout 300 python tests/tools/train_host_profile.py res18 > gpurun_out/host_prof_res18_c.log 2>&1; echo "rc=$?"
DIN_KINETO=0 timeout 300 python tests/tools/train_host_profile.py vgg16 > gpurun_out/host_prof_vgg16_c.log 2>&1; echo "rc=$?"
grep "host issue" gpurun_out/host_prof_*_c.log
