#!/bin/bash
# the multi-device paths of one process (needs >= 2 GPUs: gpurun --gpus 2): group tests and the one-process group bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q --timeout 300 -k "group or multi or device or worker" 2>&1 | tail -2
python tools/group_bench.py 512 256 2>&1 | tail -4
