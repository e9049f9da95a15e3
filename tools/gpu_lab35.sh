#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_backward_gpu.py -m gpu -q -s -k "collective_training" 2>&1 | grep -E "passed|failed|collective full|^E  " | head
timeout 300 python tests/tools/train_stage2_synthetic.py --steps 12 2>&1 | grep -v Warn | tail -2
timeout 300 python tests/tools/train_stage2_synthetic.py --steps 12 --freeze-backbone --batch 8 2>&1 | grep -v Warn | tail -1
