#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python tools/sweep_prec.py > gpurun_out/r2o_prec.log 2>&1; cat gpurun_out/r2o_prec.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
tail -4 gpurun_out/r2o_pytest.log
python tools/config3_scan.py > gpurun_out/r2o_config3_scan.txt 2>&1; tail -12 gpurun_out/r2o_config3_scan.txt
OPTSTENCIL_ROUND_LOG=gpurun_out/r2o_rounds.txt python tools/scan_timing.py config3full > gpurun_out/r2o_scan.txt 2>&1; cat gpurun_out/r2o_scan.txt
