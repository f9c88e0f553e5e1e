#!/bin/bash
# ncu --set full of the start-grid screening launch (config 3 (i)); the plain command runs first
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python tools/config3_scan.py > gpurun_out/c3.txt 2>&1; rc=$?; tail -2 gpurun_out/c3.txt | cut -c1-160
if [ $rc -eq 0 ]; then
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:objective_kernel --launch-skip 1 -c 1 -o gpurun_out/c3scan_after -f python tools/config3_scan.py > gpurun_out/c3_ncu.log 2>&1
  tail -2 gpurun_out/c3_ncu.log
fi
