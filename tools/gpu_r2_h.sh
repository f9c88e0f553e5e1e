#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python tools/sweep_prec.py > gpurun_out/r2h_prec.log 2>&1; cat gpurun_out/r2h_prec.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -4 gpurun_out/r2h_pytest.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"
head -c 250 gpurun_out/r2h_bench.json; echo
