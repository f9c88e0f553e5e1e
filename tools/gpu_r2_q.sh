#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" = "1" ]; then
  time python bench.py --steps 5 --warmup 3 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?"
  tail -4 gpurun_out/r2q_bench.err
  head -c 300 gpurun_out/r2q_bench.json; echo
  timeout 600 python -m pytest tests/test_bench_contract.py -m gpu -q --timeout 600 2>&1 | tail -3
else
  time python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $NG --steps 5 --warmup 3 > gpurun_out/r2q_bench_n$NG.json 2> gpurun_out/r2q_bench_n$NG.err; echo "bench rc=$?"
  tail -4 gpurun_out/r2q_bench_n$NG.err
  head -c 300 gpurun_out/r2q_bench_n$NG.json; echo
fi
