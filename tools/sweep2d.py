"""2-D throughput of the batched objective (the reference's default dimensionality)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import optimize_stencil_b200 as sb  # noqa: E402
from optimize_stencil_b200.extended_stencil import StencilSymmetric2D  # noqa: E402

st = StencilSymmetric2D()
rng = np.random.default_rng(0)
for N, B in ((50, 81), (50, 4096), (200, 4096), (200, 65536), (2000, 256), (8192, 16)):
    x = np.linspace(0, np.pi, N)
    ctx = sb.Context(2, [np.cos(x)] * 2, [0.5 * (1. - np.cos(x))] * 2, [x, x])
    P = np.array([0.6857978707687641, 0.10963493621568149, -0.12476172846832859, -0.12479339137808687]) \
        + 1e-3 * rng.standard_normal((B, 4))
    P[:, 0] -= 2e-3
    co = st.coefficients_batch(P)
    ctx.objective_batch(co[:8])
    best = 1e30
    for _ in range(3):
        t0 = time.perf_counter()
        out = ctx.objective_batch(co)
        wall = (time.perf_counter() - t0) * 1e3
        best = min(best, ctx.last_batch_ms or wall)
    print("2D N=%5d B=%6d: %9.3f ms  %8.2f G evals/s (%s)  nan=%d" % (
        N, B, best, B * float(N) ** 2 / best / 1e6, "kernel events" if ctx.last_batch_ms else "wall incl. Python", int((out[3] > 0).sum())))
    ctx.close()
