#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tests/tools/inv3_split_study.py 2>&1 | grep -E "^split" | tee gpurun_out/inv3_split_study.log
