#!/bin/bash
# GPU lab run 1: conv / stem / pool kernels vs torch references. Each pytest node in its own process so a
# trapped kernel (watchdog) cannot poison the following cases.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/lab1_smi.txt 2>&1
python - <<'PY' > gpurun_out/lab1_ids.txt 2>gpurun_out/lab1_collect.err
import subprocess, sys
out = subprocess.run([sys.executable, "-m", "pytest", "tests/test_conv_gpu.py", "--collect-only", "-q", "-m", "gpu"],
                     capture_output=True, text=True).stdout
for line in out.splitlines():
    if "::" in line:
        print(line.strip())
PY
: > gpurun_out/lab1.log
while IFS= read -r id; do
  echo "=== $id" >> gpurun_out/lab1.log
  timeout 300 python -m pytest "$id" -q -x -m gpu -p no:cacheprovider 2>&1 | tail -25 >> gpurun_out/lab1.log
  echo "rc=$?" >> gpurun_out/lab1.log
done < gpurun_out/lab1_ids.txt
grep -E "^===|passed|failed|error|Error|err " gpurun_out/lab1.log | tail -60
