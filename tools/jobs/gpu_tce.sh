#!/bin/bash
O=gpurun_out/tce; mkdir -p $O
timeout 900 python -m pytest tests/test_backward_gpu.py tests/test_e2e_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -rA -s -k "tce" > $O/pytest.log 2>&1; echo "rc=$?"
grep -E "passed|failed|error" $O/pytest.log | tail -3; grep -E "^FAILED|^ERROR|^E  |\[tce full|isolated:|logits max" $O/pytest.log | head -60
