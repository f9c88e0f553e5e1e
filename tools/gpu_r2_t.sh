#!/bin/bash
# round-2 GPU call T (8 GPUs): after fixes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q --timeout 300 -k "group or devices or worker" > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_pytest.log
tail -3 gpurun_out/r2t_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $NG --steps 5 --warmup 3 > gpurun_out/r2t_bench_n$NG.json 2> gpurun_out/r2t_bench_n$NG.err; echo "bench rc=$?"
head -c 200 gpurun_out/r2t_bench_n$NG.json; echo
OPTSTENCIL_ROUND_LOG=gpurun_out/r2t_rounds.txt timeout 600 python tools/scan_timing.py config3full > gpurun_out/r2t_scan_timing.txt 2>&1; cat gpurun_out/r2t_scan_timing.txt
