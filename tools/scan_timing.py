"""Optimize.optimize() end to end in its three execution modes -- one search after the other, lock-stepped in this
process, lock-stepped in worker processes (one GPU per worker, round robin) -- with a check that every mode prints
exactly the same lines.  Run on the GPU box:

    python tools/scan_timing.py [config1] [config3] [config3full] [free2d]      > profiles/rNN_scan_timing.txt
"""
import contextlib
import hashlib
import io
import os
import sys
import time
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import optimize_stencil_b200 as sb  # noqa: E402
from optimize_stencil_b200 import extended_stencil as es  # noqa: E402

BASE = dict(N=3, scan_dt=True, Ngrid_low=50, Ngrid_high=200, dim=2, div_free=False, symmetric_axes=0,
            symmetric_beta=True, Y=1.0, Z=1.0, dt_multiplier=1.0, deltaxrange=[-1, 0.25], deltayrange=[-1, 0.25],
            deltazrange=[-1, 0.25], betaxyrange=[-1, 1], betaxzrange=[-1, 1], betayxrange=[-1, 1],
            betayzrange=[-1, 1], betazxrange=[-1, 1], betazyrange=[-1, 1], dtrange=[0.1, 1], weight="equal",
            weight_params=[], singlecore=False)
CASES = {
    "config1": ("BASELINE.json configs[0]: optimize_stencil.py defaults (2-D, 81 starts, 50^2 -> 200^2)", {},
                ["sequential", "lockstep"]),
    "free2d": ("2-D free stencil, 243 starts (--no-symmetric-beta)", dict(symmetric_beta=False), ["sequential", "lockstep"]),
    "config3": ("BASELINE.json configs[2] reduced: --div-free --dim 3 --dtrange 0.6718 1.0 --N 16 on 128^3 (256 searches)",
                dict(dim=3, div_free=True, dtrange=[0.6718, 1.0], N=16, Ngrid_low=128, Ngrid_high=128),
                ["sequential", "lockstep", "workers"]),
    "config3full": ("BASELINE.json configs[2]: --div-free --dim 3 --dtrange 0.6718 1.0 --N 64 on 128^3 (4096 searches)",
                    dict(dim=3, div_free=True, dtrange=[0.6718, 1.0], N=64, Ngrid_low=128, Ngrid_high=128),
                    (["workers"] if os.environ.get("OPTSTENCIL_ROUND_LOG") else ["lockstep", "workers"]) + (["sequential"] if os.environ.get("SCAN_TIMING_SEQUENTIAL") else [])),
}


def run(over, mode):
    es.Optimize.lockstep = mode != "sequential"
    args = dict(BASE, **over)
    args["workers"] = None if mode == "workers" else 0
    if mode == "workers" and os.environ.get("SCAN_TIMING_WORKERS"):
        args["workers"] = int(os.environ["SCAN_TIMING_WORKERS"])
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        t = time.perf_counter()
        opt = es.Optimize(types.SimpleNamespace(**args))
        if mode == "workers":
            opt.workers_min_starts = 0
        x, f = opt.optimize()
        el = time.perf_counter() - t
    return el, x, f, buf.getvalue(), opt


if __name__ == "__main__":
    names = sys.argv[1:] or ["config1", "config3"]
    print("devices visible: %d, host cores: %d" % (sb.device_count(), os.cpu_count()))
    run(dict(N=1), "sequential")                                      # CUDA context, module load
    for name in names:
        title, over, modes = CASES[name]
        print("== " + title)
        ref = None
        for mode in modes:
            el, x, f, text, opt = run(over, mode)
            digest = hashlib.sha1(text.encode()).hexdigest()[:12]
            if ref is None:
                ref = digest
            st = opt.dispersion.stats
            print("  %-10s %8.3f s   launches %6d  evaluated %7d  scan %s   norm=%r   output sha1 %s %s"
                  % (mode, el, st["launches"], st["evaluated"], opt.scan_stats, f, digest,
                     "" if digest == ref else "<-- DIFFERENT OUTPUT"), flush=True)
