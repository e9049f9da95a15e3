#!/bin/bash
# compute-sanitizer memcheck over the tiny end-to-end forward + training step (smoke) and the small backward cases
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_smoke.log 2>&1
echo "smoke rc=$?"; grep -E "smoke:|ERROR SUMMARY|Invalid|out of bounds|misaligned" gpurun_out/sanitizer_smoke.log | head -12
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_conv_bwd_gpu.py tests/test_backward_gpu.py tests/test_stage1_loss_gpu.py -m gpu -q -x -k "relu_pool or roi_align or (wgrad and 16) or (stem_wgrad and tensor) or dynamic_infer_bwd or readout_bwd or layernorm_bwd or ce_metrics or mean_axis or collective_geometry" > gpurun_out/sanitizer_tests.log 2>&1
echo "tests rc=$?"; grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds|misaligned" gpurun_out/sanitizer_tests.log | head -12
