#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
bash tools/sweep_variants.sh 512 512 "4,0,0" main in0 in6d10 main in0
} > gpurun_out/r2g_sweep.log 2>&1
cat gpurun_out/r2g_sweep.log
