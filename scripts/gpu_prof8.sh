#!/bin/bash
# ncu capture of the kernel_mode-8 kernels on C2 (one launch each) -> gpurun_out/mode8.ncu-rep
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(spread|interp)_lean' -c 2 \
  -o gpurun_out/mode8 -f python scripts/prof_c2.py ${1:-C2} 1 ${2:-0} 2>&1 | tail -5
