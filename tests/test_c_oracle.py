"""The plain-C oracle (oracle/oracle_dispersion.c) against the fixtures of the unmodified reference and
against the NumPy oracle."""
import numpy as np

import c_oracle
import np_oracle as O
from conftest import relerr


def test_c_oracle_against_reference_fixtures(golden):
    checked = 0
    for i in range(golden.n):
        c = golden.case(i)
        if np.any(np.isnan(c["params"])) or c["N"] > 128:
            continue
        g = c["grid"]
        dim = c["dim"]
        grid = O.Grid(dim, c["N"], c["Y"], c["Z"])
        w = O.weight(grid, c["wkind"], c["wparams"])
        co = c_oracle.COracle(dim, g["cos"], g["s2"], g["kcomp"], c["Y"], c["Z"], None if c["wkind"] == "equal" else w)
        S, lo, hi, nan = co.objective(c["coeffs"])
        tag = "case %d %s" % (i, c["tag"])
        assert lo[0] == c["Amin"] and hi[0] == c["Amax"], tag            # same operation order: bitwise
        if not c["gated"]:
            assert nan[0] == 0 and relerr(S[0], c["S"]) <= 1e-12, (tag, S[0], c["S"])
            om = co.export(c["coeffs"], 0)
            idx = golden.sample_indices(i, om.size)
            on_cfl = c["dt_ok"] < 1e-9   # libm asin vs numpy's: a 1-ulp difference is amplified at s -> 1
            assert np.max(relerr(om.ravel()[idx], c["omega_samples"])) <= (1e-13 if not on_cfl else 1e-7), tag
            checked += 1
    assert checked >= 50


def test_c_oracle_matches_numpy_oracle_bitwise_on_sqrtarg_and_slabs():
    grid = O.Grid(3, 21, 1.3, 0.8)
    cos, s2, kc = grid.flat_tables()
    co = O.coefficients("StencilFree3D", [0.4, .02, .03, .01, .02, .03, .01, -.02, -.03, -.01])
    C = c_oracle.COracle(3, cos, s2, kc, 1.3, 0.8)
    np.testing.assert_array_equal(C.export(co, 1), np.broadcast_to(O.sqrtarg(grid, co), grid.shape))
    full = C.objective(co)
    parts = [C.objective(co, a, b) for a, b in ((0, 4), (4, 5), (5, 21))]
    assert relerr(sum(p[0][0] for p in parts), full[0][0]) <= 1e-14
    assert min(p[1][0] for p in parts) == full[1][0] and max(p[2][0] for p in parts) == full[2][0]
    ref = O.device_contract(grid, co)
    assert relerr(full[0][0], ref[0]) <= 1e-13 and full[1][0] == ref[1] and full[2][0] == ref[2]
