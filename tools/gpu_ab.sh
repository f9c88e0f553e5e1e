#!/bin/bash
# A/B of library variants built by tools/build_variants.py:   VARIANTS="main split" bash tools/gpu_ab.sh
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for v in ${VARIANTS:-main}; do
  if [ "$v" = main ]; then unset SB200_LIB; else export SB200_LIB=build/variants/libstencil_b200_$v.so; fi
  echo "== variant $v"
  python tools/sweep_prec.py
  python tools/small_bench.py seg3d 2>&1 | grep "512 segments"
done > gpurun_out/ab.log 2>&1
cat gpurun_out/ab.log
