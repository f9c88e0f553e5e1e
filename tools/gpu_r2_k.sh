#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python tools/scan_timing.py config3 config3full > gpurun_out/r2k_scan.txt 2>&1; cat gpurun_out/r2k_scan.txt
SCAN_TIMING_WORKERS=8 python tools/scan_timing.py config3full 2>&1 | grep workers | tee -a gpurun_out/r2k_scan.txt
python tools/small_bench.py > gpurun_out/r2k_small.txt 2>&1; cat gpurun_out/r2k_small.txt
python tools/small_bench.py grid2d > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:objective_kernel -s 1 -c 1 -o gpurun_out/r2k_grid2d python tools/small_bench.py grid2d > gpurun_out/r2k_ncu1.log 2>&1; echo "ncu1 rc=$?"
python tools/small_bench.py seg3d > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:objective_kernel -s 13 -c 1 -o gpurun_out/r2k_seg3d python tools/small_bench.py seg3d > gpurun_out/r2k_ncu2.log 2>&1; echo "ncu2 rc=$?"
