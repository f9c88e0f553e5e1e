"""Developer sweep: kernel-only throughput of the batched objective for several tilings.

usage: python tools/sweep.py [N] [B]    (writes one line per configuration)
Times come from CUDA events recorded by the library around the kernel (sb200_last_batch_ms).
"""
import sys
import os
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import optimize_stencil_b200 as sb  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
configs = [tuple(int(v) for v in a.split(",")) for a in sys.argv[3:]] or \
    [(4, 0, 0), (2, 0, 0), (8, 0, 0), (1, 0, 0), (4, 256, 16), (4, 128, 64), (4, 64, 64), (8, 256, 64)]

F_ALG = 61.0
x = np.linspace(0, np.pi, N)
cos = np.cos(x)
s2 = 0.5 * (1. - cos)
rng = np.random.default_rng(0)
xc = np.array([0.5, .02, .03, .01, .02, .03, .01, -.02, -.03, -.01])
if os.environ.get("SWEEP_CASE") == "lehe":   # near-CFL stencil: most points in the HIGH class
    dlx = 0.25 * (1 - 1 / 0.96 ** 2 * np.sin(np.pi * 0.96 / 2.0))
    xc = np.array([0.96, .125, .125, .125, .125, .125, .125, dlx, 0.0, 0.0])
    P = xc + 1e-5 * rng.standard_normal((B, 10))
    P[:, 0] = 0.96 - 1e-4 * rng.random(B)
else:
    P = xc + 1e-3 * rng.standard_normal((B, 10))
co = np.zeros((B, 13))
co[:, 0] = P[:, 0]
co[:, 4:10] = P[:, 1:7]
co[:, 10:13] = P[:, 7:10]
co[:, 1] = 1.0 - 2.0 * co[:, 4] - 2.0 * co[:, 5] - 3.0 * co[:, 10]
co[:, 2] = 1.0 - 2.0 * co[:, 6] - 2.0 * co[:, 7] - 3.0 * co[:, 11]
co[:, 3] = 1.0 - 2.0 * co[:, 8] - 2.0 * co[:, 9] - 3.0 * co[:, 12]

peak = sb.fp64_peak_tflops(0, 0.5)
print("measured DFMA peak %.2f TFLOP/s (nominal 37.2)" % peak, flush=True)
YY, ZZ = float(os.environ.get("SWEEP_Y", 1.0)), float(os.environ.get("SWEEP_Z", 1.0))
ctx = sb.Context(3, [cos] * 3, [s2] * 3, [x, x / YY, x / ZZ], Y=YY, Z=ZZ)
if os.environ.get("SWEEP_W") == "plane":
    ctx.set_weight(np.exp(-(x[None, :, None] / 0.3) ** 2 - (x[None, None, :] / 0.3) ** 2))
elif os.environ.get("SWEEP_W") == "dense":
    ctx.set_weight(np.random.default_rng(1).random((N, N, N)))
ref = None
for (c, ty, tx) in configs:
    ctx.set_tiling(c, ty, tx)
    ctx.objective_batch(co[:8])  # warm-up
    best = 1e30
    for rep in range(3):
        t0 = time.perf_counter()
        out = ctx.objective_batch(co)
        wall_ms = (time.perf_counter() - t0) * 1e3
        best = min(best, ctx.last_batch_ms or wall_ms)   # the zero-copy latency path records no events
    if ref is None:
        ref = out
    dev = float(np.max(np.abs(out[0] - ref[0]) / np.abs(ref[0])))
    evals = B * float(N) ** 3 / (best * 1e-3)
    print("C=%d ty=%3d tx=%3d  %9.3f ms  %8.2f Gevals/s  %6.2f TFLOP/s(alg)  %5.1f%% of nominal  %5.1f%% of measured  maxdev %.1e nan %d"
          % (c, ty, tx, best, evals / 1e9, evals * F_ALG / 1e12, evals * F_ALG / 37.2e12 * 100,
             evals * F_ALG / (peak * 1e12) * 100, dev, int(out[3].sum())), flush=True)
