#!/bin/bash
# full ncu capture of every stem / conv launch of ONE step (2 clips = 20 frames: chunks of 16 + 4), kernel v10
# (CTA-pair kernel + resident weights on the narrow-N layers)
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none -k regex:"conv_igemm|stem_tc" -s 78 -c 28 -f -o gpurun_out/prof_step_v10 \
   python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-train-step --clips-per-gpu 2 > gpurun_out/ncu_step_v10.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_step_v10.log | cut -c1-300
ls -la gpurun_out/prof_step_v10.ncu-rep
