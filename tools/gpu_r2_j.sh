#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
{
python tools/scan_timing.py config3full
SCAN_TIMING_WORKERS=8 python tools/scan_timing.py config3full | grep workers
SCAN_TIMING_WORKERS=4 python tools/scan_timing.py config3full | grep workers
} > gpurun_out/r2j_scan.txt 2>&1
cat gpurun_out/r2j_scan.txt
