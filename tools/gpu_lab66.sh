#!/bin/bash
mkdir -p gpurun_out
timeout 80 python -m pytest tests/test_backward_gpu.py tests/test_stage1_loss_gpu.py -m gpu -q -s -x -k "(full_training_step_with_backbone and res18) or resnet18" > gpurun_out/pytest_66.log 2>&1
echo "rc=$?"; grep -E "passed|failed|^FAILED|^E  |\[full step\] loss|\[stage1 res18" gpurun_out/pytest_66.log | cut -c1-220 | head -10
grep -E "\[full step\] backbone.*bn[12]?\.weight|\[full step\] backbone.features.1.weight|downsample.1.weight" gpurun_out/pytest_66.log | sort -k5 -g -r | head -6 | cut -c1-150
