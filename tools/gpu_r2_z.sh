#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python tools/config3_scan.py > gpurun_out/r2z_c3.txt 2>&1; rc=$?; tail -2 gpurun_out/r2z_c3.txt | cut -c1-200
if [ $rc -eq 0 ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:objective_kernel --launch-skip 1 -c 1 -o gpurun_out/r2z_c3scan -f python tools/config3_scan.py > gpurun_out/r2z_ncu.log 2>&1
  tail -3 gpurun_out/r2z_ncu.log
fi
