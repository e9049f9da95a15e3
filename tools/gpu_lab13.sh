#!/bin/bash
# new tests first (no -x: collect every failure), then the rest of the suite
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_ingest_gpu.py tests/test_stage1_loss_gpu.py tests/test_backward_gpu.py -m gpu -q -s 2>&1 | tail -400 > gpurun_out/pytest_new_13.log
echo "new rc=${PIPESTATUS[0]}"; grep -E "passed|failed" gpurun_out/pytest_new_13.log | tail -3
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_ingest_gpu.py --deselect tests/test_stage1_loss_gpu.py --deselect tests/test_backward_gpu.py 2>&1 | tail -30 > gpurun_out/pytest_old_13.log
echo "old rc=${PIPESTATUS[0]}"; tail -3 gpurun_out/pytest_old_13.log
