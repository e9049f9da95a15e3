#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_backward_gpu.py -m gpu -q -s -k "training_step" > gpurun_out/pytest_new_15.log 2>&1
echo "new rc=$?"; grep -E "passed|failed" gpurun_out/pytest_new_15.log | tail -3
grep -E "^\[step .*(isolated:|logits max)" gpurun_out/pytest_new_15.log | cut -c1-170
