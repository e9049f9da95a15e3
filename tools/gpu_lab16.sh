#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_bwd_gpu.py -m gpu -q -s > gpurun_out/pytest_16.log 2>&1
echo "rc=$?"; grep -E "passed|failed|\[wgrad\]|^E  " gpurun_out/pytest_16.log | head -40
