#!/bin/bash
# Inception-v3 backward: kernel tests, full training step vs the oracle / reference fixture
O=gpurun_out/r2u; mkdir -p $O
timeout 900 python -m pytest tests/test_conv_bwd_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -rA -k "general or inception or pad0 or relu_slice" > $O/pytest_k.log 2>&1; echo "kernels rc=$?"
grep -E "passed|failed|error" $O/pytest_k.log | tail -3; grep -E "^FAILED|^ERROR|^E  " $O/pytest_k.log | head -30
timeout 900 python -m pytest tests/test_backward_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -rA -s -k "inv3_full or optimizer_loop" > $O/pytest_m.log 2>&1; echo "model rc=$?"
grep -E "passed|failed|error" $O/pytest_m.log | tail -3; grep -E "^FAILED|^ERROR|^E  |worst" $O/pytest_m.log | head -30
