"""Timing of Optimize.optimize() with and without worker processes (run as a FILE: spawn re-imports __main__)."""
import contextlib
import io
import os
import sys
import time
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optimize_stencil_b200 import extended_stencil as es  # noqa: E402

if __name__ == "__main__":
    base = dict(N=6, scan_dt=True, Ngrid_low=48, Ngrid_high=48, dim=3, div_free=True, symmetric_axes=0,
                symmetric_beta=True, Y=1.0, Z=1.0, dt_multiplier=1.0, deltaxrange=[-1, 0.25], deltayrange=[-1, 0.25],
                deltazrange=[-1, 0.25], betaxyrange=[-1, 1], betaxzrange=[-1, 1], betayxrange=[-1, 1],
                betayzrange=[-1, 1], betazxrange=[-1, 1], betazyrange=[-1, 1], dtrange=[0.3, 0.6], weight="equal",
                weight_params=[], singlecore=False)
    if len(sys.argv) > 1 and sys.argv[1] == "config3":      # 4096 (dt, delta) starts on 128^3, 2 free parameters
        base.update(N=64, Ngrid_low=128, Ngrid_high=128, dtrange=[0.6718, 1.0])
    for w in (0, 8, 16):
        a = types.SimpleNamespace(workers=w, **base)
        with contextlib.redirect_stdout(io.StringIO()):
            opt = es.Optimize(a)
            t = time.perf_counter()
            x, f = opt.optimize()
            el = time.perf_counter() - t
        print("3D div-free %d^3, N=%d, workers=%d: %.2f s  x=%s norm=%r" % (base["Ngrid_low"], base["N"], w, el, x, f), flush=True)
