#!/bin/bash
O=gpurun_out/r2d; mkdir -p $O
timeout 120 python tools/debug_fused.py 1 16 16 > $O/dbg_16.log 2>&1; echo "rc=$?"; tail -8 $O/dbg_16.log
timeout 120 python tools/debug_fused.py 2 48 64 > $O/dbg_48.log 2>&1; echo "rc=$?"; tail -8 $O/dbg_48.log
timeout 300 compute-sanitizer --tool memcheck python tools/debug_fused.py 1 16 16 > $O/sanitizer.log 2>&1; echo "sanitizer rc=$?"; grep -v "^$" $O/sanitizer.log | head -40
