#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_stage1_loss_gpu.py tests/test_backward_gpu.py -m gpu -q -s -k "stage1 or per_clip or basenet" > gpurun_out/pytest_25.log 2>&1
echo "rc=$?"; grep -E "passed|failed|^FAILED|^E  |\[stage1 step" gpurun_out/pytest_25.log | cut -c1-220 | head -30
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
