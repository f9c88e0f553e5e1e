#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python tools/small_bench.py > gpurun_out/r2z_small.txt 2>&1; cat gpurun_out/r2z_small.txt
