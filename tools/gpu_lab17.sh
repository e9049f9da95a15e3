#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_bwd_gpu.py tests/test_backward_gpu.py tests/test_edge_cases_gpu.py tests/test_head_gpu.py -m gpu -q -s -k "relu_pool or dgrad or stem_wgrad or roi_align or grad_to_f16 or full_training or optimizer or abi_reports or fixture" > gpurun_out/pytest_17.log 2>&1
echo "rc=$?"; grep -E "passed|failed|^FAILED|^E  |\[full step\]" gpurun_out/pytest_17.log | cut -c1-200 | head -80
