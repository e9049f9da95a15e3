#!/bin/bash
# round-1 session-2 call 1: whole GPU suite (incl. u8 ingest, stage-1, loss) + default bench line
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu_12.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -8 gpurun_out/pytest_gpu_12.log
timeout 600 python bench.py > gpurun_out/bench_12.json 2> gpurun_out/bench_12.err
echo "bench rc=$?"; cut -c1-3000 gpurun_out/bench_12.json; tail -3 gpurun_out/bench_12.err
