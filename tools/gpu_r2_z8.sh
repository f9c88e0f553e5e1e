#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q --timeout 300 -k "group or multi or device or worker" > gpurun_out/r2z8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z8_pytest.log
tail -3 gpurun_out/r2z8_pytest.log
OPTSTENCIL_TRACE=1 OPTSTENCIL_ROUND_LOG=gpurun_out/r2z8_rounds.txt python tools/scan_timing.py config3full > gpurun_out/r2z8_scan.txt 2>&1; cat gpurun_out/r2z8_scan.txt | cut -c1-330
A="--div-free --dim 3 --dtrange 0.6718 1.0 --N 64 --Ngrid_low 128 --Ngrid_high 128"
( time OPTSTENCIL_TRACE=1 python bin/optimize_stencil.py $A > gpurun_out/r2z8_cli_out.txt ) > gpurun_out/r2z8_cli_trace.txt 2>&1
cat gpurun_out/r2z8_cli_trace.txt | cut -c1-250
