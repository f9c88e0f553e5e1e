"""Empirical check of the error bound used by csrc/objective_fa.cuh: |2dt^2 A(regrouped) - 2dt^2 A(reference order)|
relative to 2dt^2 * Mbound, with a plain-rounding NumPy emulation of the regrouped form (the kernel uses FMAs, which
round less often).  CPU only.   python tools/regroup_error_bound.py
"""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))
import numpy as np
import np_oracle as O
rng = np.random.default_rng(3)
worst = 0.0
for trial in range(400):
    N = 24
    Y, Z = (1.0, 1.0) if trial % 2 else (float(np.exp(rng.uniform(-1, 1))), float(np.exp(rng.uniform(-1, 1))))
    g = O.Grid(3, N, Y, Z)
    scale = rng.choice([0.05, 0.3, 1.0, 10.0])
    p = rng.uniform(-0.4, 0.4, 10) * scale
    p[0] = rng.uniform(0.05, 1.2)
    co = O.coefficients("StencilFree3D", p)
    A = np.broadcast_to(O.sqrtarg(g, co), g.shape)
    dt, ax, ay, az, bxy, bxz, byx, byz, bzx, bzy, dlx, dly, dlz = co
    cx, cy, cz = [c for c in g.cos]
    sx2, sy2, sz2 = g.s2
    rY, rZ = 1.0 / (Y * 1.0) ** 2, 1.0 / (Z * 1.0) ** 2
    c2 = 2 * dt * dt
    opx, opy, opz = 1 + 2 * cx, 1 + 2 * cy, 1 + 2 * cz
    # regrouped (plain roundings; the kernel uses FMAs, which round less often)
    axy = (2 * bxy) * cy + (dlx * opx + ax)
    ayx = (2 * byx) * cx + (dly * opy + ay)
    sy2r, sz2r = sy2 * rY, sz2 * rZ
    R = c2 * (axy * sx2 + ayx * sy2r)
    bzys = c2 * ((2 * bzy) * cy)
    tzx = (2 * bxz) * cz
    tzzx = (2 * bzx) * cx + (dlz * opz + az)
    tzys = c2 * ((2 * byz) * cz)
    Ts = c2 * (tzx * sx2 + tzzx * sz2r)
    W = tzys * sy2r + (bzys * sz2r + (R + Ts))
    mb = (abs(ax) + 3 * abs(dlx) + 2 * abs(bxy) + 2 * abs(bxz)) + (abs(ay) + 3 * abs(dly) + 2 * abs(byx) + 2 * abs(byz)) * rY + (abs(az) + 3 * abs(dlz) + 2 * abs(bzx) + 2 * abs(bzy)) * rZ
    err = np.max(np.abs(W - c2 * A)) / (c2 * mb)
    worst = max(worst, err)
print('worst |A\' - A| / Mbound = %.3e = 2^%.1f (rigorous: 22u = 2^-48.5; guards sized for 2^-47 = %.3e)' % (worst, np.log2(worst), 2.0 ** -47))
