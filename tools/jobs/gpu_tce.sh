#!/bin/bash
# Dynamic_TCE_volleyball training: gradient tests, the reference's unmodified TCE script, train-step timing
O=gpurun_out/tce; mkdir -p $O
timeout 1200 python -m pytest tests/test_backward_gpu.py tests/test_e2e_gpu.py tests/test_dropin_gpu.py -m gpu -q --timeout 900 -p no:cacheprovider -rA -s -k "tce or unmodified" > $O/pytest.log 2>&1; echo "rc=$?"
grep -E "passed|failed|error" $O/pytest.log | tail -3; grep -E "^FAILED|^ERROR|^E  " $O/pytest.log | head -30
