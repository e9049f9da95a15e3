#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu_53.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -3 gpurun_out/pytest_gpu_53.log
DIN_KINETO=0 timeout 300 python tests/tools/train_host_profile.py res18 > gpurun_out/host_prof_res18_b.log 2>&1; echo "rc=$?"
DIN_KINETO=0 timeout 300 python tests/tools/train_host_profile.py vgg16 > gpurun_out/host_prof_vgg16_b.log 2>&1; echo "rc=$?"
grep "host issue" gpurun_out/host_prof_*_b.log
