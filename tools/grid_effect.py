"""The same candidates (jittered around the config-3 optimum, near the CFL limit) on grids of different size: how
much of the small-grid slowdown is the grid (lanes of a warp span a larger part of k-space: more class mixing) and how
much the launch shape.

    python tools/grid_effect.py [ncu]       (ncu: only the 128^3 launch, for a capture)
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import optimize_stencil_b200 as sb  # noqa: E402
from optimize_stencil_b200.extended_stencil import StencilIsotropicDivFree3D  # noqa: E402

rng = np.random.default_rng(0)
only = sys.argv[1:] == ["ncu"]
for centre, label in (((0.6718, -0.0135), "near the CFL limit"), ((0.45, -0.0135), "dt = 0.45")):
    for N, B in ((128, 1536), (256, 192), (512, 24), (512, 96)):
        if only and (N != 128 or centre[0] != 0.6718):
            continue
        x = np.linspace(0, np.pi, N)
        ctx = sb.Context(3, [np.cos(x)] * 3, [0.5 * (1 - np.cos(x))] * 3, [x] * 3)
        P = np.array(centre) + 1e-3 * rng.standard_normal((B, 2))
        if only:
            P[:, 0] = centre[0] - np.abs(P[:, 0] - centre[0]) - 2e-3      # all below the CFL limit: live iterates
        co = StencilIsotropicDivFree3D().coefficients_batch(P)
        for name, kw in (("precise", dict(fast=False)), ("fast", dict(fast=True))):
            if only and name == "precise":
                continue
            ctx.objective_batch(co, **kw)
            best = 1e30
            for _ in range(1 if only else 4):
                ctx.objective_batch(co, **kw)
                best = min(best, ctx.last_batch_ms * 1e-3)
            print("%-20s %4d^3 x %5d rows (%s): %8.3f ms  %7.1f G evals/s" % (label, N, B, name, 1e3 * best, B * N ** 3 / best / 1e9), flush=True)
        ctx.close()
