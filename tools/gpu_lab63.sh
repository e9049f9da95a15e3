#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_stage1_loss_gpu.py -m gpu -q -s -k "resnet18 or stage1_training_step" > gpurun_out/pytest_63.log 2>&1
echo "rc=$?"; grep -E "passed|failed|^FAILED|^E  |\[stage1" gpurun_out/pytest_63.log | cut -c1-250 | head -12
