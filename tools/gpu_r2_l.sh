#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python tools/sweep_prec.py > gpurun_out/r2l_prec.log 2>&1; cat gpurun_out/r2l_prec.log
python tools/small_bench.py > gpurun_out/r2l_small.txt 2>&1; cat gpurun_out/r2l_small.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
tail -4 gpurun_out/r2l_pytest.log
python tools/scan_timing.py config1 config3full > gpurun_out/r2l_scan.txt 2>&1; cat gpurun_out/r2l_scan.txt
