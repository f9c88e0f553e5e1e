"""BASELINE.json configs[2]: `optimize_stencil.py --div-free --dtrange 0.6718 1.0` in 3-D on a 128^3 k-grid
with a 64-step dt sweep.  Variant (i) of SURVEY.md 8d: ONE batched launch over the whole start grid
(64 dt values x 64 deltax starts = 4096 candidates of StencilIsotropicDivFree3D), gated candidates
included, checked against the oracle looped over a sample of the same points; plus the reference-API
view (norm / stencil_ok / dt_ok arrays) through Dispersion3D.evaluate_batch.

    python tools/config3_scan.py > profiles/rNN_config3_scan.txt
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import np_oracle as O  # noqa: E402
from optimize_stencil_b200.extended_stencil import Dispersion3D, StencilIsotropicDivFree3D  # noqa: E402

N, NS = 128, 64
st = StencilIsotropicDivFree3D()
d = Dispersion3D(1.0, 1.0, N, st)
dts = np.linspace(0.6718, 1.0, NS + 2)[1:-1]          # optimize.py:160 with N = 64
dls = np.linspace(-1, 0.25, NS + 2)[1:-1]             # optimize.py:155
P = np.array([(a, b) for a in dts for b in dls])
d.evaluate_batch(P[:8])                                # warm-up: context, tables
d._cache.clear()
t0 = time.perf_counter()
out = d.evaluate_batch(P)
wall = time.perf_counter() - t0
kern_ms = d._context().last_batch_ms
usable = out["norm"] < 1e6
print("config 3 (i): %d candidates x %d^3 grid in one launch: kernel %.3f ms (%.1f G evals/s), "
      "evaluate_batch wall %.1f ms; %d of %d starts feasible (the 2-D CFL range 0.6718..1.0 lies above the 3-D limit "
      "0.577 of this family for most starts)" % (len(P), N, kern_ms, len(P) * N ** 3 / kern_ms / 1e6, wall * 1e3,
                                                  int(usable.sum()), len(P)))
grid = O.Grid(3, N)
rng = np.random.default_rng(0)
sample = np.unique(np.concatenate([rng.integers(0, len(P), 24), np.flatnonzero(usable)[:8]]))
t0 = time.perf_counter()
worst = 0.0
for i in sample:
    r = O.evaluate(grid, O.coefficients("StencilIsotropicDivFree3D", P[i]))
    assert out["stencil_ok"][i] == r["stencil_ok"] and (out["dt_ok"][i] == r["dt_ok"]), i
    assert (out["norm"][i] == 1e6) == (r["omega"] is None), i
    if r["omega"] is not None:
        worst = max(worst, abs(out["norm"][i] - r["norm"]) / r["norm"])
cpu = (time.perf_counter() - t0) / len(sample)
print("oracle on %d sampled starts: stencil_ok / dt_ok bitwise equal, gating identical, max rel norm error %.2e; "
      "one core needs %.3f s per start -> %.0f s for the whole grid (%.0fx the GPU wall time)"
      % (len(sample), worst, cpu, cpu * len(P), cpu * len(P) / wall))
