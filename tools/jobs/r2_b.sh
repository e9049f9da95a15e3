#!/bin/bash
# round-2 GPU job B: precision study of the degenerate shapes, drop-in run, the tests that failed in job A, stand-alone DPI autograd
O=gpurun_out/r2b; mkdir -p $O
timeout 900 python tests/tools/edge_precision_study.py > $O/edge_precision.md 2> $O/edge_precision.err; echo "study rc=$?"; cat $O/edge_precision.md
timeout 900 python -m pytest tests/test_dropin_gpu.py tests/test_head_gpu.py -m gpu -q --timeout 900 -p no:cacheprovider -rA > $O/pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed" $O/pytest.log | tail -3; grep -E "^FAILED|^ERROR|Error" $O/pytest.log | head -20
for v in "1 1" "0 0"; do set -- $v; DIN_SMALL_EXACT=$1 DIN_SMALL_EMBED_F32=$2 timeout 600 python -m pytest tests/test_backward_gpu.py -m gpu -q -k fixture --timeout 600 -p no:cacheprovider > $O/pytest_fixture_$1$2.log 2>&1; echo "fixture exact=$1 embed=$2 rc=$?"; grep -E "passed|failed" $O/pytest_fixture_$1$2.log | tail -1; done
