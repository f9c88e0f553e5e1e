"""Tiny end-to-end exercise of every kernel variant (for compute-sanitizer runs)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import np_oracle as O  # noqa: E402
import optimize_stencil_b200 as sb  # noqa: E402

rng = np.random.default_rng(0)
for dim, N, Y, Z in ((3, 37, 1.0, 1.0), (3, 20, 1.3, 0.7), (2, 50, 1.0, 1.0), (2, 33, 2.0, 1.0), (3, 1, 1.0, 1.0)):
    fam = "StencilFree3D" if dim == 3 else "StencilFree2D"
    if N == 1:
        x = np.array([0.0])
        grid = O.Grid.from_tables(dim, [np.cos(x)] * dim, [0.5 * (1. - np.cos(x))] * dim, [x] * dim, Y, Z)
    else:
        grid = O.Grid(dim, N, Y, Z)
    cos, s2, kc = grid.flat_tables()
    with sb.Context(dim, cos, s2, kc, Y=Y, Z=Z) as ctx:
        for B in (1, 2, 3, 5, 9, 70):
            P = 0.08 * rng.standard_normal((B, len(O.parameter_names(fam))))
            P[:, 0] = rng.uniform(0.1, 1.2, B)
            co = np.stack([O.coefficients(fam, p) for p in P])
            for wk in ("unit", "plane", "dense"):
                if wk == "unit":
                    ctx.set_weight_unit()
                elif wk == "plane":
                    ctx.set_weight(rng.random((1,) + grid.shape[1:]) if dim == 3 else rng.random(grid.shape))
                else:
                    ctx.set_weight(rng.random(grid.shape))
                for cfg in ((0, 0, 0), (1, 0, 0), (2, 32, 3), (4, 0, 1), (8, 64, 0)):
                    ctx.set_tiling(*cfg)
                    ctx.objective_batch(co)
                    if N > 4 and dim == 3:
                        ctx.objective_batch(co, 1, N - 2)
        for what in (0, 1, 2):
            ctx.export(co[0], what)
print("sanity_small: done")
