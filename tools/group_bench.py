"""The device group behind the drop-in API: Dispersion3D.norm_batch(params) in ONE process over all visible GPUs
(sb200_group_*: one host thread, one context and stream per GPU) against the same call on one GPU.

    python tools/group_bench.py [B] [N]          default 4096 x 512^3 (BASELINE.json configs[3])
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import optimize_stencil_b200 as sb  # noqa: E402
from optimize_stencil_b200.extended_stencil import Dispersion3D, StencilFree3D  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 512
XC = np.array([0.5, .02, .03, .01, .02, .03, .01, -.02, -.03, -.01])
P = XC + 1e-3 * np.random.default_rng(0).standard_normal((B, 10))
G = sb.device_count()
print("devices visible: %d" % G)
res = {}
for devs in ([0], list(range(G))):
    d = Dispersion3D(1.0, 1.0, N, StencilFree3D())
    d.devices = devs
    t0 = time.perf_counter()
    d.norm_batch(P[:8 * len(devs)])                       # contexts, tables, staging buffers
    t_create = time.perf_counter() - t0
    best = 1e30
    for rep in range(3):
        t0 = time.perf_counter()
        out = d.norm_batch(P)
        best = min(best, time.perf_counter() - t0)
    res[len(devs)] = out
    print("norm_batch(%d x %d^3) on %d device(s): %8.1f ms  %8.1f G evals/s  (shard %s; first call incl. context creation %.2f s)"
          % (B, N, len(devs), 1e3 * best, B * float(N) ** 3 / best / 1e9, d._context().last_shard, t_create), flush=True)
if G > 1:
    a, b = res[1], res[G]
    print("max relative difference of the norms, %d devices vs 1: %.2e" % (G, float(np.max(np.abs(a - b) / np.abs(a)))))
    # k-slab sharding of one candidate on a large grid through the same call
    N5 = int(os.environ.get("GROUP_BENCH_N5", 2048))
    d = Dispersion3D(1.0, 1.0, N5, StencilFree3D())
    d.devices = list(range(G))
    ctx = d._context()
    C = StencilFree3D().coefficients_batch(XC[None, :])
    for shard in ("none", "slabs", "slabs-contiguous"):
        ctx.objective_batch(C, shard=shard, fast=True)
        best = 1e30
        for rep in range(5):
            t0 = time.perf_counter()
            r = ctx.objective_batch(C, shard=shard, fast=True)
            best = min(best, time.perf_counter() - t0)
        print("one candidate on %d^3, shard=%-16s %7.2f ms  %8.1f G evals/s  S=%r" % (N5, shard, 1e3 * best, float(N5) ** 3 / best / 1e9, r[0][0]), flush=True)
