#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_conv_gpu.py -m gpu -q -x -k "45 and 128" > gpurun_out/pytest_30a.log 2>&1
echo "first rc=$?"; tail -5 gpurun_out/pytest_30a.log | cut -c1-300
nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv | tail -1
timeout 300 python -m pytest tests/test_conv_gpu.py tests/test_properties_gpu.py -m gpu -q > gpurun_out/pytest_30b.log 2>&1
echo "conv rc=$?"; tail -5 gpurun_out/pytest_30b.log | cut -c1-300
