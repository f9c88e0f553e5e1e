#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
OPTSTENCIL_ROUND_LOG=gpurun_out/r2n_rounds.txt python tools/scan_timing.py config3full 2>&1 | grep -v "^  lockstep" > gpurun_out/r2n_scan.txt; cat gpurun_out/r2n_scan.txt
