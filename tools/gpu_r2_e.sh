#!/bin/bash
# round-2 GPU call E (1 GPU): full test-suite on the new kernel, scan timings, bench line, ncu launch list + full capture
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -4 gpurun_out/r2e_pytest.log
timeout 900 python tools/scan_timing.py config1 free2d config3 config3full > gpurun_out/r2e_scan_timing.txt 2>&1
cat gpurun_out/r2e_scan_timing.txt
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"
head -c 300 gpurun_out/r2e_bench.json; echo
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-parity"
$CMD > gpurun_out/r2e_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2e_launches.csv $CMD > gpurun_out/r2e_ncu_l.log 2>&1
echo "ncu launches rc=$?"
$CMD > gpurun_out/r2e_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:objective_kernel -s 3 -c 1 -o gpurun_out/r2e_fa_bench $CMD > gpurun_out/r2e_ncu_f.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out/r2e_*
