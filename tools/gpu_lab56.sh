#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_bwd_gpu.py tests/test_backward_gpu.py tests/test_edge_cases_gpu.py -m gpu -q -s -k "batch_stat or batch_statistics or optimizer_loop or edge or loud or fallback" > gpurun_out/pytest_56.log 2>&1
echo "rc=$?"; grep -E "passed|failed|^FAILED|^E  |worst rel|logits max" gpurun_out/pytest_56.log | cut -c1-250 | head -30
