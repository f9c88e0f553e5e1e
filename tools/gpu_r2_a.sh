#!/bin/bash
# round-2 GPU call A (1 GPU): tests, bench line, peak shapes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc >> gpurun_out/r2a_smi.txt; free -g >> gpurun_out/r2a_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
SB200_PEAK_VERBOSE=1 timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2a_bench.err
head -c 1500 gpurun_out/r2a_bench.json
