#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/stem_ab.py 2>&1 | tee gpurun_out/stem_ab.log | tail -8
