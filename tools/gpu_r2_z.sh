#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_pytest.log
tail -4 gpurun_out/r2z_pytest.log
