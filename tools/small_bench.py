"""Launch-size-bound cases: small 2-D grids and small segments (what SLSQP rounds look like).

    python tools/small_bench.py [case]      cases: grid2d  seg3d  single3d   (default: all)
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import optimize_stencil_b200 as sb  # noqa: E402
from optimize_stencil_b200.extended_stencil import StencilSymmetric2D, StencilIsotropicDivFree3D  # noqa: E402

cases = sys.argv[1:] or ["grid2d", "seg3d", "single3d"]
rng = np.random.default_rng(0)


def timed(ctx, co, reps=5, **kw):
    ctx.objective_batch(co, **kw)
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        ctx.objective_batch(co, **kw)
        wall = time.perf_counter() - t0
        best = min(best, (ctx.last_batch_ms * 1e-3) or wall)
    return best


if "grid2d" in cases:
    for N, B in ((50, 4096), (50, 65536), (200, 4096)):
        x = np.linspace(0, np.pi, N)
        ctx = sb.Context(2, [np.cos(x)] * 2, [0.5 * (1 - np.cos(x))] * 2, [x, x])
        P = np.array([0.6856, 0.1096, -0.1248, -0.1248]) + 1e-3 * rng.standard_normal((B, 4))
        co = StencilSymmetric2D().coefficients_batch(P)
        for name, kw in (("precise", dict(fast=False)), ("fast", dict(fast=True))):
            t = timed(ctx, co, **kw)
            print("2-D %d^2 x %6d candidates (%s): %8.3f ms  %7.1f G evals/s" % (N, B, name, 1e3 * t, B * N * N / t / 1e9), flush=True)
if "seg3d" in cases:
    N = 128
    x = np.linspace(0, np.pi, N)
    ctx = sb.Context(3, [np.cos(x)] * 3, [0.5 * (1 - np.cos(x))] * 3, [x] * 3)
    # live iterates: 2 dt^2 max(sqrtarg) = 1.89 (the CFL limit of delta_x = -0.0135 is dt = 0.5356); hopeless ones: the
    # same stencil at the lower end of config 3's dt range, 2 dt^2 max(sqrtarg) = 3.15 -- the extrema-only sweep
    for (dt0, what) in ((0.52, "live"), (0.6718, "hopeless")):
        for nseg in (64, 512, 4096):
            P = np.array([dt0, -0.0135]) + 1e-3 * rng.standard_normal((3 * nseg, 2))
            co = StencilIsotropicDivFree3D().coefficients_batch(P)
            t = timed(ctx, co, seg_size=3, fast=False)
            print("3-D 128^3, %5d segments of 3 (an SLSQP round of %d %s searches): %8.3f ms  %7.1f G evals/s"
                  % (nseg, nseg, what, 1e3 * t, 3 * nseg * N ** 3 / t / 1e9), flush=True)
            t = timed(ctx, co, fast=False)
            print("3-D 128^3, the same %5d rows as ONE batch:                     %8.3f ms  %7.1f G evals/s"
                  % (3 * nseg, 1e3 * t, 3 * nseg * N ** 3 / t / 1e9), flush=True)
if "single3d" in cases:
    N = 128
    x = np.linspace(0, np.pi, N)
    ctx = sb.Context(3, [np.cos(x)] * 3, [0.5 * (1 - np.cos(x))] * 3, [x] * 3)
    co = StencilIsotropicDivFree3D().coefficients_batch(np.array([0.52, -0.0135]) + 1e-3 * rng.standard_normal((3, 2)))
    ctx.objective_batch(co)
    t0 = time.perf_counter()
    for _ in range(200):
        ctx.objective_batch(co)
    t = (time.perf_counter() - t0) / 200
    print("3-D 128^3, one stand-alone launch of 3 rows (host call, launch + sync): %7.1f us  %6.1f G evals/s" % (1e6 * t, 3 * N ** 3 / t / 1e9))
