"""BASELINE.json configs[4]: single-candidate objective on a 2048^3 k-grid, k-slab-sharded across the
ranks of a torchrun job, combined with one all-gather (sharding.py).  Verifies every rank's slab
against the slab-wise NumPy oracle on a few thin sub-slabs and prints timing.

    torchrun --nproc-per-node G tools/config5_slabs.py [N]
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import np_oracle as O  # noqa: E402
import optimize_stencil_b200 as sb  # noqa: E402
from optimize_stencil_b200 import sharding  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
x = np.linspace(0, np.pi, N)
cos, s2 = np.cos(x), 0.5 * (1. - np.cos(x))
ctx = sb.Context(3, [cos] * 3, [s2] * 3, [x] * 3, device=local)
xc = np.array([0.5, .02, .03, .01, .02, .03, .01, -.02, -.03, -.01])
co = O.coefficients("StencilFree3D", xc).reshape(1, -1)
sh = sharding.ShardedObjective(sharding.device_evaluator(ctx), nx=N, mode="slabs")
c_dev = torch.from_numpy(co).to(dev)
for _ in range(2):
    out = sh.evaluate(c_dev)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
K = 5
for _ in range(K):
    out = sh.evaluate(c_dev)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = (time.perf_counter() - t0) / K
res = out.cpu().numpy()[:, 0]
x0, x1, xs = sh.my_planes()            # this rank's planes: x0, x0 + xs, ... below x1
# parity of this rank's share on two thin sub-lists of its planes
ok = True
for a, b in ((x0, min(x0 + 2 * xs, x1)), (max(x0 + ((x1 - 1 - x0) // xs - 1) * xs, x0), x1)):
    got = ctx.objective_batch(co, a, b, xs)
    Sr, amin, amax, nn = O.device_contract(O.Grid(3, N, xslab=(a, b, xs)), co[0])
    ok &= abs(got[0][0] - Sr) <= 1e-12 * abs(Sr) and got[2][0] == max(amax, 0.0)
flag = torch.tensor([1.0 if ok else 0.0], device=dev)
if world > 1:
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("config 5: N=%d^3 single candidate, %d rank(s), slabs of %d planes: %.3f ms per objective, %.1f G evals/s, "
          "S=%.15g Amax=%.17g, slab parity vs oracle: %s" % (N, world, len(range(x0, x1, xs)), dt * 1e3, float(N) ** 3 / dt / 1e9,
                                                             res[0], res[2], "ok" if flag.item() == 1.0 else "FAILED"))
if world > 1:
    dist.destroy_process_group()
