#!/bin/bash
# round-2 GPU call C (1 GPU): kernel variant sweep
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
{
BLOCKS="256 128" bash tools/sweep_variants.sh 512 512 "4,0,0 8,0,0" main deg9 vimax deg9vimax
echo "#### near-CFL batch"
SWEEP_CASE=lehe bash tools/sweep_variants.sh 512 512 "4,0,0 8,0,0" main deg9vimax
echo "#### parity of deg9vimax"
SB200_LIB=build/variants/libstencil_b200_deg9vimax.so python tools/parity_report.py
} > gpurun_out/r2c_sweep.log 2>&1
tail -60 gpurun_out/r2c_sweep.log
