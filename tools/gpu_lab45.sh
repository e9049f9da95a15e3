#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/pytest_gpu_45.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -2 gpurun_out/pytest_gpu_45.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_backward_gpu.py tests/test_conv_bwd_gpu.py tests/test_stage1_loss_gpu.py -m gpu -q -x -k "dynamic_infer_bwd or layernorm_bwd or readout_bwd or (stem_wgrad and tensor and False) or ce_metrics or relu_pool" > gpurun_out/racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/racecheck.log | head -8
