#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_backward_gpu.py -m gpu -q -s -k "full_training_step_with_backbone and collective" > gpurun_out/pytest_59.log 2>&1
echo "rc=$?"; grep -E "passed|failed|^FAILED|^E  |worst|loss" gpurun_out/pytest_59.log | cut -c1-250 | head -20
grep "full step\] " gpurun_out/pytest_59.log | sort -k5 -g -r | head -8
