"""TEST / BENCH INFRASTRUCTURE ONLY -- the UNMODIFIED reference evaluated slab by slab.

`Dispersion3D.norm` of the reference (dispersion.py:235-255) needs about seven N^3-sized temporaries: 7 GiB per
worker at 512^3, and 2048^3 does not fit at all.  Its broadcasting formulae work on any table lengths, though
(SURVEY.md section 7, item 1), so this module drives the reference's OWN class over x-slabs: it builds one
`Dispersion3D` object, and for every slab replaces the x-axis tables (and the dense `k` of that slab, computed with
init2's formula, dispersion.py:344-356) before calling the reference's `norm`.  Every arithmetic pass that is timed
is the reference's code; only the table set-up is ours, and it is excluded from the timing like the reference's own
init2 would be.

Caveat: the gate (`stencil_ok`, `dt_ok`) is then evaluated per slab, not on the whole grid; for candidates that are
feasible everywhere -- the benchmark's -- the sum of the slab norms is the whole-grid norm up to summation order.
Used by bench.py's CPU legs (`cpu_baseline`, `--impl reference`) and checked in tests/test_oracle_golden.py.
"""
import numpy as np

import ref_shim

_STATE = {}


def reference_state(N, slab, Y=1.0, Z=1.0, stencil="StencilFree3D"):
    key = (N, slab, Y, Z, stencil)
    st = _STATE.get(key)
    if st is None:
        es = ref_shim.load_reference()
        d = es.Dispersion3D(Y, Z, 4, getattr(es, stencil)())   # a tiny grid; every table is replaced below
        d.N = N                                                 # norm's (pi/(N-1))**dim factor, dispersion.py:253
        x = np.linspace(0, np.pi, N)
        slabs = []
        for a in range(0, N, slab):
            xs = x[a:a + slab]
            t = dict(kappax=xs.reshape(-1, 1, 1), kappay=x.reshape(1, -1, 1), kappaz=x.reshape(1, 1, -1))
            t["kx"], t["ky"], t["kz"] = t["kappax"], t["kappay"] / Y, t["kappaz"] / Z
            t["kappamesh"] = [t["kappax"], t["kappay"], t["kappaz"]]
            t["kmesh"] = (t["kx"], t["ky"], t["kz"])
            t["k"] = np.sqrt((t["kappax"]) ** 2 + (t["kappay"] / Y) ** 2 + (t["kappaz"] / Z) ** 2)
            for ax in "xyz":
                t["coskappa" + ax] = np.cos(t["kappa" + ax])
                t["s" + ax + "2"] = 0.5 * (1. - t["coskappa" + ax])
            slabs.append(t)
        st = _STATE[key] = (d, slabs)
    return st


def norm_slabwise(params, N, slab, **kw):
    """Sum over x-slabs of the reference's own Dispersion3D.norm(params) (each already scaled by (pi/(N-1))^3)."""
    d, slabs = reference_state(N, slab, **kw)
    total = 0.0
    for t in slabs:
        d.__dict__.update(t)
        d._parameters = None      # the reference's one-entry cache is keyed by the parameters alone
        total += float(d.norm(params))
    return total
