"""Parity report (run on the GPU box): max relative errors of the CUDA path against the NumPy oracle.

usage: python tools/parity_report.py > profiles/rNN_parity_report.txt
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import np_oracle as O  # noqa: E402
import optimize_stencil_b200 as sb  # noqa: E402


def relerr(a, b):
    with np.errstate(invalid="ignore", divide="ignore"):
        e = np.abs(a - b) / np.abs(b)
    return np.where(b == 0, np.abs(a), e)


cases = [
    ("Lehe dt=0.96, 3D 256^3 (config 2)", "StencilSymmetric3D", 3, 256,
     [0.96, .125, .125, .125, 0.25 * (1 - 1 / 0.96 ** 2 * np.sin(np.pi * 0.96 / 2.0)), 0, 0]),
    ("Lehe dt=0.96, 2D 200^2", "StencilSymmetric2D", 2, 200,
     [0.96, 0.125, 0.25 * (1 - 1 / 0.96 ** 2 * np.sin(np.pi * 0.96 / 2.0)), 0.0]),
    ("div-free optimum on the CFL limit, 2D 200^2", "StencilIsotropicDivFree2D", 2, 200,
     [0.6718000000000016, -0.013484133612260314]),
    ("default optimum, 2D 200^2", "StencilSymmetric2D", 2, 200,
     [0.6857978707687641, 0.10963493621568149, -0.12476172846832859, -0.12479339137808687]),
    ("config 4 centre, 3D 128^3", "StencilFree3D", 3, 128, [0.5, .02, .03, .01, .02, .03, .01, -.02, -.03, -.01]),
    ("Yee at 0.999 of CFL, 3D 96^3", "StencilSymmetric3D", 3, 96, [0.999 / np.sqrt(3), 0, 0, 0, 0, 0, 0]),
    ("Yee at dt = 0.2, 3D 96^3", "StencilSymmetric3D", 3, 96, [0.2, 0, 0, 0, 0, 0, 0]),
]
print("%-46s %12s %12s %10s %10s %s" % ("case", "max rel omega", "rel S", "Amin bits", "Amax bits", "class share low/mid/high/exact"))
for name, fam, dim, N, p in cases:
    grid = O.Grid(dim, N)
    co = O.coefficients(fam, p)
    ref = O.evaluate(grid, co)
    cos, s2, kc = grid.flat_tables()
    with sb.Context(dim, cos, s2, kc) as ctx:
        S, Amin, Amax, nan = ctx.objective_batch(co)
        om = ctx.export(co, 0)[0]
    s2v = co[0] ** 2 * O.sqrtarg(grid, co)
    share = [np.mean(s2v <= .25), np.mean((s2v > .25) & (s2v <= .75)), np.mean((s2v > .75) & (s2v < 1 - 2 ** -10)),
             np.mean(s2v >= 1 - 2 ** -10)]
    print("%-46s %12.3e %12.3e %10s %10s %s" % (
        name, np.max(relerr(om, ref["omega"])), abs(S[0] - ref["S"]) / abs(ref["S"]),
        "equal" if Amin[0] == min(ref["Amin"], 0.0) else "DIFF", "equal" if Amax[0] == ref["Amax"] else "DIFF",
        "/".join("%.3f" % v for v in share)))
