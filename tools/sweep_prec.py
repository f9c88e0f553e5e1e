"""fast vs precise series on the 512 x 512^3 batch (kernel-event times)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import optimize_stencil_b200 as sb
N, B = 512, 512
x = np.linspace(0, np.pi, N); cos = np.cos(x); s2 = 0.5 * (1. - cos)
rng = np.random.default_rng(0)
xc = np.array([0.5, .02, .03, .01, .02, .03, .01, -.02, -.03, -.01])
from optimize_stencil_b200.extended_stencil import StencilFree3D
co = StencilFree3D().coefficients_batch(xc + 1e-3 * rng.standard_normal((B, 10)))
ctx = sb.Context(3, [cos] * 3, [s2] * 3, [x] * 3)
for name, kw in (("fast", dict(fast=True)), ("precise", dict(fast=False)), ("screening", dict(scan=True))):
    ctx.objective_batch(co[:8], **kw)
    best = min((ctx.objective_batch(co, **kw), ctx.last_batch_ms)[1] for _ in range(3))
    print("%-9s %8.3f ms  %7.2f G evals/s" % (name, best, B * float(N) ** 3 / best / 1e6), flush=True)
