#!/bin/bash
O=gpurun_out/r2h; mkdir -p $O
timeout 300 python -m pytest tests/test_conv_gpu.py -m gpu -q -k stem_pair --timeout 300 -p no:cacheprovider -s > $O/pytest_a.log 2>&1; rc=$?; echo "pair tests rc=$rc"; grep -E "stem pair|passed|failed|AssertionError:" $O/pytest_a.log | head -20
