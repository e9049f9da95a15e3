#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke
timeout 600 python -m pytest tests/test_backward_gpu.py tests/test_conv_bwd_gpu.py -m gpu -q -s -k "reduces_loss or stem_wgrad or relu_pool" 2>&1 | grep -E "passed|failed|overfit|^E  " | head
