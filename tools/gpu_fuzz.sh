#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python tools/fuzz_long.py ${CASES:-4000} > gpurun_out/fuzz_long.txt 2>&1; tail -3 gpurun_out/fuzz_long.txt
python tools/parity_report.py > gpurun_out/parity_report.txt 2>&1; tail -12 gpurun_out/parity_report.txt
