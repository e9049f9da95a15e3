#!/bin/bash
# GPU lab run 2: head kernels + end-to-end parity, one pytest node per process.
mkdir -p gpurun_out
: > gpurun_out/lab2.log
for f in tests/test_head_gpu.py tests/test_e2e_gpu.py; do
python - "$f" <<'PY' > gpurun_out/lab2_ids.txt 2>gpurun_out/lab2_collect.err
import subprocess, sys
out = subprocess.run([sys.executable, "-m", "pytest", sys.argv[1], "--collect-only", "-q", "-m", "gpu"],
                     capture_output=True, text=True).stdout
for line in out.splitlines():
    if "::" in line:
        print(line.strip())
PY
while IFS= read -r id; do
  echo "=== $id" >> gpurun_out/lab2.log
  timeout 600 python -m pytest "$id" -q -x -s -m gpu -p no:cacheprovider 2>&1 | grep -v "Warning\|warnings.warn" | tail -30 >> gpurun_out/lab2.log
done < gpurun_out/lab2_ids.txt
done
grep -E "^===|passed|failed|rror|\[e2e\]|assert" gpurun_out/lab2.log | tail -80
