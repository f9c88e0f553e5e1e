#!/bin/bash
# after a change of the extrema-only sweep: its parity tests, the hopeless-rounds timing, the start-grid launch
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q -x --timeout 200 -k "hopeless or screening or segments" 2>&1 | tail -1
python tools/small_bench.py seg3d 2>&1 | grep "hopeless\|the same" | tail -4
python tools/config3_scan.py 2>&1 | tail -2 | cut -c1-130
