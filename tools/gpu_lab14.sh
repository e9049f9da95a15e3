#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_ingest_gpu.py tests/test_stage1_loss_gpu.py tests/test_backward_gpu.py tests/test_edge_cases_gpu.py -m gpu -q -s -k "model_u8 or basenet or checkpoint or training_step or optimizer or abi_reports" > gpurun_out/pytest_new_14.log 2>&1
echo "new rc=$?"; grep -E "passed|failed" gpurun_out/pytest_new_14.log | tail -3
