#!/bin/bash
O=gpurun_out/r2q; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv1_fused" -c 1 -o $O/ncu_fused -f python bench.py --no-cpu-baseline --no-train-step --no-e2e --steps 1 --warmup 3 > $O/ncu_fused.log 2>&1; echo "ncu rc=$?"
