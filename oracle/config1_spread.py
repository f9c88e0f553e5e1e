"""TEST INFRASTRUCTURE ONLY -- how far does the REFERENCE's own default optimisation move when its objective changes
in the last bit?

BASELINE.json asks that optimised coefficients "agree with the reference's to SciPy's tolerance".  SLSQP stops when the
objective changes by less than ftol = 1e-6 between iterates, and the default problem (config 1: StencilSymmetric2D,
81 starts, N = 50 -> 200) has a flat valley along the CFL constraint, so runs whose objective values differ in the
16th digit can stop at visibly different points.  This script measures that on the unmodified reference (imported
through oracle/ref_shim.py): Dispersion.norm is replaced, from outside the reference tree, by a restatement of its ten
lines (dispersion.py:235-255) whose only change is the final reduction `F = f.sum()`:

    baseline    f.sum()                      (NumPy pairwise summation, the reference itself)
    up / down   nextafter(f.sum(), +-inf)    (one ulp)
    reversed    f.ravel()[::-1].sum()        (same terms, other pairing)
    fsum        math.fsum(f.ravel())         (correctly rounded sum)
    jitter      f.sum() * (1 + u), u uniform in +-2^-52, seeded

and prints the optimum of each run and its distance from the baseline.  tests/test_dropin_api.py takes its
coefficient tolerance for config 1 from the spread recorded in profiles/r02_config1_spread.txt.

    python oracle/config1_spread.py > profiles/r02_config1_spread.txt        (build container: needs the reference)
"""
import contextlib
import io
import math
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402


def make_norm(variant):
    rng = np.random.default_rng(0)

    def norm(self, parameters):           # dispersion.py:235-255, the reduction swapped as the docstring says
        if np.any(np.isnan(parameters)):
            return 1e6
        omega = self.omega(parameters)
        if omega is None:
            return 1e6
        f = self.w * (omega - self.k) ** 2
        if variant == "baseline":
            F = f.sum()
        elif variant == "up":
            F = np.nextafter(f.sum(), np.inf)
        elif variant == "down":
            F = np.nextafter(f.sum(), -np.inf)
        elif variant == "reversed":
            F = np.ascontiguousarray(f.ravel()[::-1]).sum()
        elif variant == "fsum":
            F = np.float64(math.fsum(f.ravel()))
        else:
            F = f.sum() * (1.0 + rng.uniform(-1, 1) * 2.0 ** -52)
        F = F * (np.pi / (self.N - 1)) ** self.dim * self.Y ** 2 * self.Z ** 2
        return F
    return norm


def main():
    es = ref_shim.load_reference()
    disp_mod = sys.modules[es.__name__ + ".dispersion"]
    args = types.SimpleNamespace(N=3, scan_dt=True, Ngrid_low=50, Ngrid_high=200, dim=2, div_free=False,
                                 symmetric_axes=0, symmetric_beta=True, Y=1.0, Z=1.0, dt_multiplier=1.0,
                                 deltaxrange=[-1, 0.25], deltayrange=[-1, 0.25], deltazrange=[-1, 0.25],
                                 betaxyrange=[-1, 1], betaxzrange=[-1, 1], betayxrange=[-1, 1], betayzrange=[-1, 1],
                                 betazxrange=[-1, 1], betazyrange=[-1, 1], dtrange=[0.1, 1], weight="equal",
                                 weight_params=[], singlecore=True)
    original = disp_mod.Dispersion.norm
    results = {}
    print("reference: %s (unmodified; Dispersion.norm's final f.sum() perturbed from outside)" % ref_shim.REFERENCE_ROOT)
    print("%-9s %-22s %-22s %-22s %-22s %-22s" % ("variant", "dt", "betaxy", "deltax", "deltay", "norm"))
    try:
        for variant in ("baseline", "up", "down", "reversed", "fsum", "jitter"):
            disp_mod.Dispersion.norm = make_norm(variant)
            with contextlib.redirect_stdout(io.StringIO()):
                x, fmin = es.Optimize(args).optimize()
            results[variant] = (np.asarray(x, dtype=float), float(fmin))
            print("%-9s %s %-22r" % (variant, " ".join("%-22r" % float(v) for v in x), float(fmin)), flush=True)
    finally:
        disp_mod.Dispersion.norm = original
    x0, f0 = results["baseline"]
    print()
    print("distance from the baseline run:")
    worst_x, worst_f = 0.0, 0.0
    for variant, (x, f) in results.items():
        dx, df = float(np.max(np.abs(x - x0))), abs(f - f0)
        worst_x, worst_f = max(worst_x, dx), max(worst_f, df)
        print("%-9s max |coefficient difference| %.3e   |norm difference| %.3e" % (variant, dx, df))
    print()
    print("spread: coefficients %.3e, norm %.3e   (SLSQP ftol = 1e-6 on the norm)" % (worst_x, worst_f))


if __name__ == "__main__":
    main()
