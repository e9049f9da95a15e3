#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/tools/train_host_profile.py res18 > gpurun_out/host_prof_res18.log 2>&1; echo "rc=$?"
timeout 300 python tests/tools/train_host_profile.py vgg16 > gpurun_out/host_prof_vgg16.log 2>&1; echo "rc=$?"
grep "host issue" gpurun_out/host_prof_*.log
