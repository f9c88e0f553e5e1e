#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python tools/debug_search.py divfree3d_small gpurun_out/rows_gpu_main.npy
SB200_LIB=build/variants/libstencil_b200_in0.so python tools/debug_search.py divfree3d_small gpurun_out/rows_gpu_in0.npy
SB200_LIB=build/variants/libstencil_b200_in6d10.so python tools/debug_search.py divfree3d_small gpurun_out/rows_gpu_in6d10.npy
