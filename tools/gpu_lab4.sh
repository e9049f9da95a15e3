#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/probe.log
for v in 0 1 2 3 4; do
  DIN_CONV_VARIANT=$v timeout 300 python tools/probe_conv.py >> gpurun_out/probe.log 2>&1
  echo "rc=$?" >> gpurun_out/probe.log
done
cat gpurun_out/probe.log | grep -v Warn
