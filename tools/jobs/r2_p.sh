#!/bin/bash
O=gpurun_out/r2p; mkdir -p $O
timeout 900 python -m pytest tests/test_e2e_gpu.py tests/test_fullsize_gpu.py -m gpu -q -k "tce" --timeout 600 -p no:cacheprovider -s > $O/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|\[e2e\]" $O/pytest.log | tail -8
