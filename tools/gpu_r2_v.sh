#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
{
for i in 1 2 3; do
echo "blas threads limited to 1:"; python tools/scan_timing.py config1 2>&1 | grep -E "lockstep"
echo "blas threads untouched:";   OPTSTENCIL_BLAS_THREADS=0 python tools/scan_timing.py config1 2>&1 | grep -E "lockstep"
done
} > gpurun_out/r2v_c1.txt
cut -c1-120 gpurun_out/r2v_c1.txt
