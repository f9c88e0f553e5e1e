"""The NumPy oracle against the fixtures produced by the UNMODIFIED reference (oracle/make_golden.py).

This is what pins the oracle: SURVEY.md 8(c) -- no reference test pins a numeric value of norm or
omega, so parity is anchored on outputs of the reference itself run in the build container.
"""
import numpy as np

import np_oracle as O
from conftest import relerr


def _grid(c):
    return O.Grid(c["dim"], c["N"], c["Y"], c["Z"])


def test_case_count(golden):
    assert golden.n >= 80


def test_tables_bitwise(golden):
    """dispersion.py:268-281 / 339-357: the oracle's tables equal the reference's, bit for bit
    (same NumPy build; on another CPU np.cos may differ in the last bit, hence <= 1 ulp there)."""
    for g in golden.grids:
        dim, N, Y, Z = int(g["key"][0]), int(g["key"][1]), g["key"][2], g["key"][3]
        cos, s2, kc = O.Grid(dim, N, Y, Z).flat_tables()
        for a in range(dim):
            assert np.max(np.abs(cos[a] - g["cos"][a])) <= 2.3e-16
            assert np.max(np.abs(s2[a] - g["s2"][a])) <= 1.2e-16
            np.testing.assert_array_equal(kc[a], g["kcomp"][a])  # linspace / Y: no libm involved


def test_coefficients_bitwise(golden):
    """stencil.py:121-129 and every _fill_coefficients chain."""
    seen = set()
    for i in range(golden.n):
        c = golden.case(i)
        if np.any(np.isnan(c["params"])):
            continue
        np.testing.assert_array_equal(O.coefficients(c["stencil"], c["params"]), c["coeffs"])
        seen.add(c["stencil"])
    assert seen == set(O.STENCILS)


def test_objective_against_reference_outputs(golden):
    for i in range(golden.n):
        c = golden.case(i)
        grid = _grid(c)
        if np.any(np.isnan(c["params"])):
            r = O.evaluate(grid, np.full(7, np.nan))
            assert r["norm"] == 1e6 == c["norm"] and r["stencil_ok"] == -1 == c["stencil_ok"] and c["dt_ok"] == -1
            continue
        co = O.coefficients(c["stencil"], c["params"])
        w = O.weight(grid, c["wkind"], c["wparams"])
        r = O.evaluate(grid, co, w, c["dtmul"])
        tag = "case %d %s" % (i, c["tag"])
        # constraint values steer SLSQP: bitwise on the same NumPy, <= a few ulp across CPUs
        assert relerr(r["Amin"], c["Amin"]) <= 1e-14, tag
        assert relerr(r["Amax"], c["Amax"]) <= 1e-14, tag
        assert relerr(r["stencil_ok"], c["stencil_ok"]) <= 1e-14, tag
        assert abs(r["dt_ok"] - c["dt_ok"]) <= 1e-14 * max(1.0, abs(c["dt_ok"])), tag
        assert (r["omega"] is None) == bool(c["gated"]), tag
        if c["gated"]:
            assert r["norm"] == 1e6 == c["norm"], tag
            continue
        assert relerr(r["norm"], c["norm"]) <= 1e-12, tag
        assert relerr(r["S"], c["S"]) <= 1e-12, tag
        idx = golden.sample_indices(i, r["omega"].size)
        on_cfl = c["dt_ok"] < 1e-9  # max s within 1e-9 of 1: d(omega)/ds is ~1e4..inf there
        tol = 1e-13 if not on_cfl else 1e-7
        assert np.max(relerr(r["omega"].ravel()[idx], c["omega_samples"])) <= tol, tag
        if c["omega_full"] is not None:
            assert np.max(relerr(r["omega"], c["omega_full"])) <= tol, tag


def test_reference_smoke_assertions(golden):
    """The three checks the reference's own tests make (tests/test_dispersion.py:25-61)."""
    by_tag = {golden.case(i)["tag"]: golden.case(i) for i in range(golden.n)}
    for tag in ("reftest-2d", "reftest-3d"):
        c = by_tag[tag]
        r = O.evaluate(_grid(c), O.coefficients(c["stencil"], c["params"]))
        assert r["stencil_ok"] > 0 and r["dt_ok"] > 0 and not np.any(np.isnan(r["omega"]))
    c = by_tag["reftest-2d-bad"]
    assert O.evaluate(_grid(c), O.coefficients(c["stencil"], c["params"]))["stencil_ok"] < 0


def test_device_contract_matches_evaluate(golden):
    for i in range(0, golden.n, 5):
        c = golden.case(i)
        if c["gated"] or np.any(np.isnan(c["params"])):
            continue
        grid = _grid(c)
        co = O.coefficients(c["stencil"], c["params"])
        w = O.weight(grid, c["wkind"], c["wparams"])
        S, Amin, Amax, nan = O.device_contract(grid, co, w)
        assert S == O.evaluate(grid, co, w, c["dtmul"])["S"] and nan == 0
        assert Amin == c["Amin"] or relerr(Amin, c["Amin"]) < 1e-14


def test_slabs_tile_the_grid():
    """x-slab evaluation: slabs partition sum / min / max."""
    c = O.coefficients("StencilFree3D", [0.5, .02, .03, .01, .02, .03, .01, -.02, -.03, -.01])
    full = O.device_contract(O.Grid(3, 24), c)
    parts = [O.device_contract(O.Grid(3, 24, xslab=(a, b)), c) for a, b in ((0, 5), (5, 16), (16, 24))]
    assert relerr(sum(p[0] for p in parts), full[0]) < 1e-13
    assert min(p[1] for p in parts) == full[1] and max(p[2] for p in parts) == full[2]


def test_slabwise_reference_driver_equals_the_reference():
    """oracle/ref_slab.py (bench.py's CPU arm: the unmodified reference's own Dispersion3D.norm over x-slabs with
    overridden tables) against the reference evaluated on the whole grid."""
    import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("the unmodified reference is not present (neither /root/reference nor baseline/_ref)")
    import ref_slab
    es = ref_shim.load_reference()
    rng = np.random.default_rng(8)
    xc = np.array([0.5, .02, .03, .01, .02, .03, .01, -.02, -.03, -.01])
    d = es.Dispersion3D(1.0, 1.0, 30, es.StencilFree3D())
    for _ in range(3):
        p = xc + 1e-3 * rng.standard_normal(10)
        whole = float(d.norm(p))
        for slab in (30, 7, 1):
            assert abs(ref_slab.norm_slabwise(p, 30, slab) - whole) <= 1e-13 * whole
