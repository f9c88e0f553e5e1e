#!/bin/bash
# final single-GPU validation: tests, smoke, bench, scan timing, ncu launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/val_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/val_pytest.log
tail -4 gpurun_out/val_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/val_smoke.log 2>&1; tail -2 gpurun_out/val_smoke.log
time python bench.py --steps 5 --warmup 3 > gpurun_out/val_bench.json 2> gpurun_out/val_bench.err; echo "bench rc=$?"
head -c 200 gpurun_out/val_bench.json; echo
python tools/scan_timing.py config3full 2>&1 | tee gpurun_out/val_scan.txt
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-parity"
$CMD > gpurun_out/val_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/val_launches.csv $CMD > gpurun_out/val_ncu_l.log 2>&1
echo "ncu launches rc=$?"
