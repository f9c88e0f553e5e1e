#!/bin/bash
# round-2 GPU call D (1 GPU): inner-series variants, export pipeline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
{
bash tools/sweep_variants.sh 512 512 "4,0,0" main in0 in7 in6d10 in0v
echo "#### near-CFL batch"
SWEEP_CASE=lehe bash tools/sweep_variants.sh 512 512 "4,0,0" main in0
echo "#### parity of main (deg 9 + inner 6)"
python tools/parity_report.py
} > gpurun_out/r2d_sweep.log 2>&1
tail -30 gpurun_out/r2d_sweep.log
{
python tools/export_bench.py 512
SB200_COPY_THREADS=1 python tools/export_bench.py 512 | grep Context.export | tail -1
SB200_COPY_THREADS=4 python tools/export_bench.py 512 | grep Context.export | tail -1
SB200_COPY_THREADS=16 python tools/export_bench.py 512 | grep Context.export | tail -1
} > gpurun_out/r2d_export.log 2>&1
cat gpurun_out/r2d_export.log
