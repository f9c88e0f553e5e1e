#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u_pytest.log
tail -5 gpurun_out/r2u_pytest.log
python tools/scan_timing.py config1 free2d config3 config3full > gpurun_out/r2u_scan.txt 2>&1; cat gpurun_out/r2u_scan.txt
