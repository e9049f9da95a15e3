#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_conv_bwd_gpu.py -m gpu -q -k "pack or bn_gamma" > gpurun_out/pytest_51.log 2>&1
echo "rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest_51.log | cut -c1-250 | head -20
DIN_KINETO=1 timeout 300 python tests/tools/train_host_profile.py res18 > gpurun_out/host_prof_res18.log 2>&1; echo "rc=$?"
DIN_KINETO=1 timeout 300 python tests/tools/train_host_profile.py vgg16 > gpurun_out/host_prof_vgg16.log 2>&1; echo "rc=$?"
grep "host issue" gpurun_out/host_prof_*.log
grep -h "pack_weight_kernel\|bn_gamma_grad_kernel\|Self CUDA time total" gpurun_out/host_prof_*.log | cut -c1-60,150-260
