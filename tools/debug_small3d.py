import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import np_oracle as O
import optimize_stencil_b200 as sb
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
grid = O.Grid(3, N)
cos, s2, kc = grid.flat_tables()
P = []
for dt in (0.7, 0.85):
    for dl in (-0.58333333, -0.16666667, 0.02):
        for h in (0.0, 1.4901161193847656e-08):
            P.append((dt + h, dl))
co = np.stack([O.coefficients("StencilIsotropicDivFree3D", p) for p in P])
with sb.Context(3, cos, s2, kc) as ctx:
    for seg in (0, 3, 1):
        S, lo, hi, nan = ctx.objective_batch(co, seg_size=seg)
        print("seg", seg)
        for b in range(len(co)):
            Sr, amin, amax, nn = O.device_contract(grid, co[b])
            flag = "" if (hi[b] == max(amax, 0.0) and lo[b] == (amin if amin < 0 else 0.0) and ((nn > 0) == (nan[b] > 0)) and (nn > 0 or abs(S[b] - Sr) <= 1e-12 * abs(Sr))) else "  <-- MISMATCH"
            print("  %-28s S %-22r ref %-22r Amax %-20r ref %-20r nan %d/%d%s" % (P[b], S[b], Sr, hi[b], amax, nan[b], nn, flag))
