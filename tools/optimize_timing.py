"""Optimize.optimize() end to end with SciPy's own finite differences vs the exact-FD `jac` callables
(optimize.py: exact_fd_jac).  Prints wall time, launch statistics and whether the two searches are
bit-identical.  Run on the GPU box:  python tools/optimize_timing.py"""
import contextlib
import io
import os
import sys
import time
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optimize_stencil_b200 import extended_stencil as es  # noqa: E402

BASE = dict(N=3, scan_dt=True, Ngrid_low=50, Ngrid_high=200, dim=2, div_free=False, symmetric_axes=0,
            symmetric_beta=True, Y=1.0, Z=1.0, dt_multiplier=1.0, deltaxrange=[-1, 0.25], deltayrange=[-1, 0.25],
            deltazrange=[-1, 0.25], betaxyrange=[-1, 1], betaxzrange=[-1, 1], betayxrange=[-1, 1],
            betayzrange=[-1, 1], betazxrange=[-1, 1], betazyrange=[-1, 1], dtrange=[0.1, 1], weight="equal",
            weight_params=[], singlecore=True)
RUNS = {"config1 default (2D, 81 starts, 50^2 -> 200^2)": {},
        "2D free stencil, 243 starts": dict(symmetric_beta=False, N=3),
        "config3-like 3D div-free 128^3, 64 dt starts": dict(dim=3, div_free=True, dtrange=[0.6718, 1.0], N=64,
                                                            Ngrid_low=128, Ngrid_high=128),
        "3D free stencil 48^3, N=1 (1 start x 10 dof)": dict(dim=3, symmetric_beta=False, N=1, Ngrid_low=48,
                                                           Ngrid_high=64, dtrange=[0.3, 0.6])}

if __name__ == "__main__":
    es.Optimize(types.SimpleNamespace(**dict(BASE, N=1))).dispersion.norm(np.array([0.5, 0.0, 0.0, 0.0]))  # warm-up
    for name, over in RUNS.items():
        out = {}
        for flag in ((True,) if os.environ.get("OPT_TIMING_JAC_ONLY") else (False, True)):
            es.Optimize.exact_fd_jac = flag
            with contextlib.redirect_stdout(io.StringIO()) as buf:
                opt = es.Optimize(types.SimpleNamespace(**dict(BASE, **over)))
                t = time.perf_counter()
                x, f = opt.optimize()
                el = time.perf_counter() - t
            out[flag] = (x, f, buf.getvalue(), el, dict(opt.dispersion.stats))
        if False not in out:
            print("%s: exact-FD jac %.3f s; launches %d, evaluated %d; norm=%r" % (
                name, out[True][3], out[True][4]["launches"], out[True][4]["evaluated"], out[True][1]), flush=True)
            continue
        same = np.array_equal(out[0][0], out[1][0]) and out[0][1] == out[1][1] and out[0][2] == out[1][2]
        print("%s: scipy-FD %.3f s, exact-FD jac %.3f s, %s; launches %d, evaluated %d; norm=%r" % (
            name, out[0][3], out[1][3], "bit-identical" if same else "DIFFERENT", out[1][4]["launches"],
            out[1][4]["evaluated"], out[1][1]), flush=True)
