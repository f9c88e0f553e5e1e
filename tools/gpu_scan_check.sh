#!/bin/bash
# single-GPU check after a kernel or scan change: GPU tests, the three series variants, config 3 (start grid and full
# scan with its per-launch log), differential fuzz.   gpurun --timeout 2400 -- 'bash tools/gpu_scan_check.sh'
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/scan_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/scan_pytest.log
tail -4 gpurun_out/scan_pytest.log
python tools/sweep_prec.py > gpurun_out/scan_prec.log 2>&1; cat gpurun_out/scan_prec.log
python tools/config3_scan.py 2>&1 | tail -2 | cut -c1-150
OPTSTENCIL_ROUND_LOG=gpurun_out/scan_rounds.txt python tools/scan_timing.py config3full > gpurun_out/scan_scan.txt 2>&1; cat gpurun_out/scan_scan.txt
python tools/fuzz_long.py 800 > gpurun_out/scan_fuzz.txt 2>&1; tail -2 gpurun_out/scan_fuzz.txt
