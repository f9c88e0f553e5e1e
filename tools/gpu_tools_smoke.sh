#!/bin/bash
# every developer tool once, at small sizes: does it still run against the current API?
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/tool_$name.log 2>&1; echo "$name rc=$? : $(tail -1 gpurun_out/tool_$name.log | cut -c1-150)"; }
run export_bench python tools/export_bench.py 128
run group_bench python tools/group_bench.py 256 128
run optimize_timing python tools/optimize_timing.py
run sanity_small python tools/sanity_small.py
run workers_timing python tools/workers_timing.py
run debug_small3d python tools/debug_small3d.py 16
run sweep python tools/sweep.py 128 64
run sweep2d python tools/sweep2d.py
run config5_slabs python tools/config5_slabs.py 256
run debug_search python tools/debug_search.py divfree gpurun_out/rows.npy
run grid_effect python tools/grid_effect.py
run regroup python tools/regroup_error_bound.py
