#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/probe_conv.py > gpurun_out/probe.log 2>&1; echo "probe rc=$?"; grep -v Warn gpurun_out/probe.log
BENCH_ARGS=--no-cpu-baseline bash tools/gpu_lab5.sh
