#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
{
for i in 1 2; do
echo "== current"; python tools/sweep_prec.py | head -2
echo "== lib739"; SB200_LIB=build/variants/libstencil_b200_lib739.so python tools/sweep_prec.py | head -2
done
} > gpurun_out/r2r_ab.log 2>&1
cat gpurun_out/r2r_ab.log
