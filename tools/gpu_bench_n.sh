#!/bin/bash
# the driver's multi-GPU launch of bench.py:   N=8 bash tools/gpu_bench_n.sh      (gpurun --gpus $N)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
N=${N:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 $BENCH_ARGS > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
tail -2 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, d['e2e']['value'], d['e2e'].get('parts'), d['roofline']['frac'], d['clocks'])
for k,v in (d.get('extra') or {}).items(): print(k, json.dumps(v)[:260])
PY
