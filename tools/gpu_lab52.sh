#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_52.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -6 gpurun_out/pytest_gpu_52.log
