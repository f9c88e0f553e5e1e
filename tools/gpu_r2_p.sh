#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
tail -12 gpurun_out/r2p_pytest.log
python tools/sweep_prec.py > gpurun_out/r2p_prec.log 2>&1; cat gpurun_out/r2p_prec.log
python tools/config3_scan.py > gpurun_out/r2p_config3_scan.txt 2>&1; tail -3 gpurun_out/r2p_config3_scan.txt | cut -c1-200
OPTSTENCIL_ROUND_LOG=gpurun_out/r2p_rounds.txt python tools/scan_timing.py config3full > gpurun_out/r2p_scan.txt 2>&1; cat gpurun_out/r2p_scan.txt
python tools/fuzz_long.py 1500 > gpurun_out/r2p_fuzz.txt 2>&1; tail -3 gpurun_out/r2p_fuzz.txt
