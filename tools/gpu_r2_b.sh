#!/bin/bash
# round-2 GPU call B (2+ GPUs): multi-device tests, bench at N = number of visible GPUs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_dropin_api.py -m gpu -x -q --timeout 600 -k "group or devices or worker or lockstep" > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 3 --warmup 3 > gpurun_out/r2b_bench_n$NG.json 2> gpurun_out/r2b_bench_n$NG.err; echo "bench rc=$?"
tail -c 800 gpurun_out/r2b_bench_n$NG.err
head -c 600 gpurun_out/r2b_bench_n$NG.json
