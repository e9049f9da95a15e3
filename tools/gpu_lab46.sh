#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_bwd_gpu.py -m gpu -q -k "maxpool3s2 or scatter2 or stride2 or bn_gamma or stem_wgrad" > gpurun_out/pytest_46.log 2>&1
echo "rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest_46.log | cut -c1-250 | head -20
