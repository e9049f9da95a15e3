#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_backward_gpu.py tests/test_edge_cases_gpu.py -m gpu -q -s -k "full_training_step or optimizer_loop or no_cpu or fallback or loud" > gpurun_out/pytest_47.log 2>&1
echo "rc=$?"; grep -E "passed|failed|^FAILED|^E  |\[full step\] loss" gpurun_out/pytest_47.log | cut -c1-250 | head -30
grep -E "\[full step\] backbone" gpurun_out/pytest_47.log | grep -v "features\.[0-9]*\.\(weight\|bias\)" | sort -t'L' -k2 | cut -c1-160 | head -80
