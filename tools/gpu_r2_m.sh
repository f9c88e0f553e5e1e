#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python tools/sweep_prec.py > gpurun_out/r2m_prec.log 2>&1; cat gpurun_out/r2m_prec.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
tail -4 gpurun_out/r2m_pytest.log
python tools/small_bench.py seg3d single3d > gpurun_out/r2m_small.txt 2>&1; cat gpurun_out/r2m_small.txt
