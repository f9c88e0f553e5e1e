#!/bin/bash
# round-2 GPU call F (1 GPU): full test-suite on the two-series kernel, scan timings with worker processes, bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -6 gpurun_out/r2f_pytest.log
timeout 900 python tools/scan_timing.py config1 config3 config3full > gpurun_out/r2f_scan_timing.txt 2>&1
cat gpurun_out/r2f_scan_timing.txt
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
head -c 300 gpurun_out/r2f_bench.json; echo
python tools/parity_report.py > gpurun_out/r2f_parity_report.txt 2>&1; tail -9 gpurun_out/r2f_parity_report.txt
