"""The drop-in `extended_stencil` package: same surface and same numbers as the reference.

Every test runs twice: on CPU with the oracle-backed context injected (tests/fake_context.py) to
check the host logic, and under `-m gpu` against the real CUDA context.  Expected
values come from the fixtures the unmodified reference produced (oracle/make_golden.py).
"""
import os
import pickle
import subprocess
import sys
import types

import numpy as np
import pytest

import fake_context
import np_oracle as O
from conftest import relerr, GOLDEN


@pytest.fixture(params=["oracle-backed", pytest.param("cuda", marks=pytest.mark.gpu)])
def es(request, monkeypatch):
    """`oracle-backed`: host logic on CPU with tests/fake_context.py injected.
    `cuda` (-m gpu): the very same assertions against the real sm_100a library."""
    if request.param == "oracle-backed":
        fake_context.install(monkeypatch)
    from optimize_stencil_b200 import extended_stencil
    return extended_stencil


def make_dispersion(es, c):
    st = getattr(es, c["stencil"])()
    d = es.Dispersion2D(c["Y"], c["N"], st) if c["dim"] == 2 else es.Dispersion3D(c["Y"], c["Z"], c["N"], st)
    d.dt_multiplier = c["dtmul"]
    if c["wkind"] != "equal":
        es.weight_functions[c["wkind"]](*c["wparams"])(d)
    return d


def check_case(es, c, tol_norm=1e-12):
    d = make_dispersion(es, c)
    p = c["params"]
    tag = "case %d %s" % (c["id"], c["tag"])
    with np.errstate(invalid="ignore", divide="ignore"):
        sok, dtok, nrm = d.stencil_ok(p), d.dt_ok(p), d.norm(p)
        om = d.omega(p)
    assert relerr(np.asarray(sok, dtype=float).ravel()[0], c["stencil_ok"]) <= 1e-14, tag
    assert abs(np.asarray(dtok, dtype=float).ravel()[0] - c["dt_ok"]) <= 1e-14 * max(1, abs(c["dt_ok"])), tag
    assert (om is None) == bool(c["gated"]), tag
    if c["gated"]:
        assert nrm == 1e6, tag
    else:
        assert relerr(nrm, c["norm"]) <= tol_norm, (tag, nrm, c["norm"])
    return d


def test_every_golden_case_through_the_reference_api(es, golden):
    for i in range(golden.n):
        c = golden.case(i)
        if c["N"] > 128 and c["dim"] == 3:
            continue  # the 256^3 case runs on the GPU (test_gpu_parity.py); too slow for the CPU suite
        check_case(es, c)


def test_surface_matches_reference_names(es):
    for name in ("Optimize", "Dispersion", "Dispersion2D", "Dispersion3D", "Stencil", "StencilFlags", "namedarray",
                 "get_stencil", "weight_functions", "minmax", "make_single_max_constraint",
                 "make_single_min_constraint", "__version__") + tuple(O.STENCILS):
        assert hasattr(es, name), name
    d = es.Dispersion3D(1.5, 2.0, 12, es.StencilFree3D())
    for attr in ("w", "dx", "Y", "Z", "dim", "N", "dt_multiplier", "stencil", "kappamesh", "kappax", "kappay", "kappaz",
                 "kx", "ky", "kz", "kmesh", "dks", "k", "coskappax", "coskappay", "coskappaz", "sx2", "sy2", "sz2",
                 "parameters", "coefficients", "init2_run"):
        assert hasattr(d, attr), attr
    g = O.Grid(3, 12, 1.5, 2.0)
    np.testing.assert_array_equal(d.k, g.k)
    np.testing.assert_array_equal(d.coskappay, g.cos[1])
    np.testing.assert_array_equal(d.sz2, g.s2[2])
    assert d.kz.shape == (1, 1, 12) and d.dks == (g.kcomp[0].ravel()[1], g.kcomp[1].ravel()[1], g.kcomp[2].ravel()[1])
    d2 = es.Dispersion2D(1, 9, es.StencilFree2D())
    assert d2.Z == 1.0 and d2.k.shape == (9, 9) and d2.w == 1.0


def test_sentinels_and_errors(es):
    d = es.Dispersion2D(1, 20, es.StencilFree2D())
    nanp = [0.5, np.nan, 0, 0, 0]
    assert d.norm(nanp) == 1e6 and d.stencil_ok(nanp) == -1 and d.dt_ok(nanp) == -1. and d.omega(nanp) is None
    assert d.norm([1.5, 0, 0, 0, 0]) == 1e6          # dt beyond CFL
    assert d.norm([0.0, 0, 0, 0, 0]) == 1e6          # dt <= 0
    assert d.norm([-.1, 0, 0, 0, 0]) == 1e6
    d.dt_multiplier = 1.3                              # gate passes with s > 1 -> NaNs in omega -> ValueError
    with pytest.raises(ValueError):
        d.norm([0.85, 0, 0, 0, 0])
    with pytest.raises(ValueError):
        d.omega([0.85, 0, 0, 0, 0])
    d.dt_multiplier = 1.0
    d.parameters = [0.3, 0.3, -0.3, 0.3, 0.7]         # tests/test_dispersion.py:39-47 (bad stencil)
    assert d.stencil_ok(d.parameters) < 0
    with pytest.raises(ValueError):
        d.sqrtres                                      # NaNs in sqrt -> ValueError (dispersion.py:118-120)
    with pytest.raises(ValueError):
        es.StencilFree2D().coefficients_batch(np.zeros((2, 3)))


def test_parameters_are_copied_and_properties_work(es):
    d = es.Dispersion3D(1, 1, 10, es.StencilDivFree3D())
    p = np.array([0.3, -0.10, -0.60, 0.2])            # tests/test_dispersion.py:50-61
    d.parameters = p
    p[0] = 9.0
    assert d.parameters[0] == 0.3
    assert d.coefficients.shape == (1,) and d.coefficients.deltaz[0] == 0.2
    g = O.Grid(3, 10)
    A = O.sqrtarg(g, O.coefficients("StencilDivFree3D", [0.3, -0.10, -0.60, 0.2]))
    np.testing.assert_array_equal(d.sqrtarg, A)
    np.testing.assert_array_equal(d.sqrtres, np.sqrt(A))
    assert d.stencil_ok(d.parameters) > 0 and d.dt_ok(d.parameters) > 0
    assert not np.any(np.isnan(d.omega(d.parameters)))


def test_one_launch_serves_objective_constraints_and_fd_points(es):
    d = es.Dispersion2D(1, 30, es.StencilSymmetric2D())
    x = np.array([0.5, 0.05, -0.05, -0.04])
    d.norm(x)
    assert d.stats["launches"] == 1 and d.stats["evaluated"] == 5
    d.stencil_ok(x), d.dt_ok(x)
    for i in range(4):                                 # scipy's forward-difference points
        xi = x + np.diag(np.full(4, 1.4901161193847656e-08))[i]
        d.norm(xi), d.stencil_ok(xi), d.dt_ok(xi)
    assert d.stats["launches"] == 1 and d.stats["hits"] == 14
    d.fd_prefetch = False
    d.norm(x + 1e-3)
    assert d.stats["launches"] == 2 and d.stats["evaluated"] == 6
    d.w = 2.0                                          # a new weight invalidates cached sums
    assert relerr(d.norm(x), 2 * es.Dispersion2D(1, 30, es.StencilSymmetric2D()).norm(x)) < 1e-15


def test_batch_api_matches_scalar_api(es):
    d = es.Dispersion3D(1.2, 0.9, 14, es.StencilFree3D())
    rng = np.random.default_rng(4)
    P = 0.08 * rng.standard_normal((12, 10))
    P[:, 0] = np.linspace(0.1, 1.2, 12)
    P[3, 4] = np.nan
    out = d.evaluate_batch(P)
    launches = d.stats["launches"]
    assert launches == 1
    ref = es.Dispersion3D(1.2, 0.9, 14, es.StencilFree3D())
    for i in range(12):
        # the constraints come from the bit-exact extrema; the sum may differ in its last bits between two
        # launches of different batch composition (the contract is determinism per launch shape, and 1e-12)
        assert relerr(out["norm"][i], ref.norm(P[i])) <= 1e-13
        assert out["stencil_ok"][i] == np.asarray(ref.stencil_ok(P[i]), dtype=float).ravel()[0]
        assert out["dt_ok"][i] == np.asarray(ref.dt_ok(P[i]), dtype=float).ravel()[0]
    np.testing.assert_array_equal(d.norm_batch(P), out["norm"])


def test_a_batch_is_a_throughput_launch_by_its_evaluations(es):
    """Few rows on a large grid (a rank's share of the benchmark batch) take the throughput path -- no cache warming, the
    library picks series and screening variant -- like many rows do; small batches warm the scalar API's cache."""
    d = es.Dispersion3D(1.0, 1.0, 12, es.StencilFree3D())
    P = 0.05 * np.random.default_rng(8).standard_normal((6, 10))
    P[:, 0] = 0.4
    small = d.evaluate_batch(P)
    assert len(d._cache) == 6
    d2 = es.Dispersion3D(1.0, 1.0, 12, es.StencilFree3D())
    d2.throughput_evals = 6 * 12 ** 3          # this very batch now counts as large
    large = d2.evaluate_batch(P)
    assert len(d2._cache) == 0 and d2.stats["launches"] == 1 and d2.stats["evaluated"] == 6
    for k in ("stencil_ok", "dt_ok"):
        np.testing.assert_array_equal(small[k], large[k])          # from the bit-exact extrema either way
    assert np.max(relerr(large["norm"], small["norm"])) <= 1e-13


def test_pickle_roundtrip_drops_the_device_handle(es):
    d = es.Dispersion2D(1.0, 25, es.StencilIsotropic2D())
    es.weight_functions["xaxis_soft"](0.4)(d)
    x = [0.4, 0.05, -0.03]
    v = d.norm(x)
    d2 = pickle.loads(pickle.dumps(d))
    assert d2._ctx is None and len(d2._cache) == 0
    assert d2.norm(x) == v


def test_weight_shapes_and_values(es):
    for dim in (2, 3):
        st = es.StencilFree2D() if dim == 2 else es.StencilFree3D()
        d = es.Dispersion2D(1.3, 11, st) if dim == 2 else es.Dispersion3D(1.3, 0.7, 11, st)
        g = O.Grid(dim, 11, 1.3, 0.7)
        es.weight_functions["xaxis"]()(d)
        np.testing.assert_array_equal(d.w, O.weight(g, "xaxis"))
        assert d.w.shape == (11,) * dim
        es.weight_functions["xaxis_soft"]()(d)
        np.testing.assert_array_equal(d.w, O.weight(g, "xaxis_soft"))
        es.weight_functions["xaxis_soft"](0.25)(d)
        np.testing.assert_array_equal(d.w, O.weight(g, "xaxis_soft", (0.25,)))
    assert set(es.weight_functions) == {"xaxis", "xaxis_soft"}


def test_get_stencil_table(es):
    """Same selections as the reference for every flag combination (reference tests/test_stencil.py)."""
    expect = {(2, 0, 0, 0): "StencilFree2D", (2, 0, 0, 1): "StencilIsotropic2D", (2, 0, 1, 0): "StencilSymmetric2D",
              (2, 0, 1, 1): "StencilIsotropic2D", (2, 1, 0, 0): "StencilDivFree2D",
              (2, 1, 0, 1): "StencilIsotropicDivFree2D", (2, 1, 1, 0): "StencilIsotropicDivFree2D",
              (2, 1, 1, 1): "StencilIsotropicDivFree2D",
              (3, 0, 0, 0): "StencilFree3D", (3, 0, 0, 1): "StencilCylinder3D", (3, 0, 0, 2): "StencilIsotropic3D",
              (3, 0, 1, 0): "StencilSymmetric3D", (3, 0, 1, 1): None, (3, 0, 1, 2): "StencilIsotropic3D",
              (3, 1, 0, 0): "StencilDivFree3D", (3, 1, 0, 1): "StencilCylinderDivFree3D",
              (3, 1, 0, 2): "StencilIsotropicDivFree3D", (3, 1, 1, 0): "StencilIsotropicDivFree3D",
              (3, 1, 1, 1): None, (3, 1, 1, 2): "StencilIsotropicDivFree3D"}
    for (dim, df, sb, sa), name in expect.items():
        a = types.SimpleNamespace(dim=dim, div_free=bool(df), symmetric_beta=bool(sb), symmetric_axes=sa)
        if name is None:
            with pytest.raises(NotImplementedError):
                es.get_stencil(a)
        else:
            assert type(es.get_stencil(a)).__name__ == name
    with pytest.raises(NotImplementedError):
        es.get_stencil(types.SimpleNamespace(dim=4, div_free=False, symmetric_beta=True, symmetric_axes=0))
    assert es.StencilIsotropicDivFree3D().dof == 2 and es.StencilFree3D().dof == 10
    assert es.StencilSymmetric2D().Parameters._fields == ["dt", "betaxy", "deltax", "deltay"]


def test_compiled_coefficient_fill_keeps_every_bit():
    """Stencil.coefficients_batch runs the reference's ordered copies and alpha formulae as a few array operations;
    it must agree bit for bit with the step-by-step procedure and with the oracle's restatement, for every family."""
    from optimize_stencil_b200 import extended_stencil as es_
    rng = np.random.default_rng(12)
    for name in sorted(O.STENCILS):
        st = getattr(es_, name)()
        P = np.concatenate([rng.standard_normal((20, st.dof)), 1e3 * rng.standard_normal((5, st.dof)), np.zeros((1, st.dof))])
        P[3, 0] = -0.0
        got = st.coefficients_batch(P)
        np.testing.assert_array_equal(got.view(np.int64), st._coefficients_batch_sequential(P).view(np.int64), err_msg=name)
        ref = np.stack([O.coefficients(name, p) for p in P])
        np.testing.assert_array_equal(got.view(np.int64), ref.view(np.int64), err_msg=name)


def default_args(**over):
    base = dict(N=3, scan_dt=True, Ngrid_low=50, Ngrid_high=200, dim=2, div_free=False, symmetric_axes=0,
                symmetric_beta=True, Y=1.0, Z=1.0, dt_multiplier=1.0, deltaxrange=[-1, 0.25],
                deltayrange=[-1, 0.25], deltazrange=[-1, 0.25], betaxyrange=[-1, 1], betaxzrange=[-1, 1],
                betayxrange=[-1, 1], betayzrange=[-1, 1], betazxrange=[-1, 1], betazyrange=[-1, 1],
                dtrange=[0.1, 1], weight="equal", weight_params=[], singlecore=True)
    base.update(over)
    return types.SimpleNamespace(**base)


OPT_RUNS = {"divfree": dict(div_free=True, dtrange=[0.6718, 1.0]),
            "xaxis_soft": dict(weight="xaxis_soft", weight_params=[0.3], N=2),
            "divfree3d_small": dict(dim=3, div_free=True, dtrange=[0.55, 1.0], Ngrid_low=16, Ngrid_high=24, N=2),
            "default": {}}


def run_optimize(es, name, capsys=None):
    z = np.load(GOLDEN + "/optimise_runs.npz")
    opt = es.Optimize(default_args(**OPT_RUNS[name]))
    assert type(opt.stencil).__name__ == str(z[name + "_stencil"])
    x, fmin = opt.optimize()
    # SLSQP stops when |f - f_prev| < ftol = 1e-6 (SciPy default, absolute): two runs whose objective
    # values differ in the 16th digit take different paths through a flat valley and stop up to a few ftol
    # apart.  The bar (BASELINE.json: "coefficients must agree ... to SciPy's tolerance") is therefore a few
    # ftol on the minimum -- and never worse than the reference's by more than 2 ftol; the coefficients then
    # agree to ~sqrt(ftol / curvature).
    ref = float(z[name + "_norm"])
    assert abs(fmin - ref) <= 3e-6 * max(1.0, abs(ref)), (name, fmin, ref)
    assert fmin <= ref + 2e-6 * max(1.0, abs(ref)), (name, fmin, ref)
    # ... and OUR objective at the REFERENCE's optimum is the reference's minimum (objective parity there)
    assert relerr(opt.dispersion_high.norm(z[name + "_x"]), ref) <= 1e-12
    # The well-conditioned runs agree to ~1e-3 in every coefficient (3.6e-4 observed in dt for the default
    # run, whose minimum then differs by 7e-7).  "xaxis_soft" weights only a narrow band around the kx axis,
    # so delta_y barely enters the objective: the valley floor is flat to 1e-6 over ~1e-2 in delta_y; dt,
    # beta_xy and delta_x are still determined to ~1e-3.
    dx_ = np.abs(x - z[name + "_x"])
    if name == "xaxis_soft":
        assert np.max(dx_[:3]) <= 2e-3 and dx_[3] <= 0.1, (name, x, z[name + "_x"])
    elif name == "default":
        # Config 1.  The bar is MEASURED, not chosen: oracle/config1_spread.py re-runs the unmodified reference with
        # the last bit of its own `f.sum()` perturbed (+-1 ulp, reversed summation order, correctly rounded sum,
        # random last-bit jitter); its optimum then moves by up to 1.5e-4 in the coefficients and 3.6e-7 in the norm
        # (profiles/r02_config1_spread.txt).  Three times that spread is "agreement to SciPy's tolerance" here.
        assert np.max(dx_) <= 3 * 1.5e-4, (name, x, z[name + "_x"])
        assert abs(fmin - ref) <= 3 * 3.636e-7, (name, fmin, ref)
    else:
        assert np.max(dx_) <= 1e-3, (name, x, z[name + "_x"])
    return opt, x, fmin, z


@pytest.mark.parametrize("name", ["divfree", "xaxis_soft", "divfree3d_small"])
def test_optimize_matches_reference_result(es, name, capsys):
    opt, x, fmin, z = run_optimize(es, name)
    out = capsys.readouterr().out
    assert out.count("x0=") >= int(z[name + "_nstarts"])   # one line per start (+ the polish line)
    assert "Using stencil " + str(z[name + "_stencil"]) in out
    launches = opt.dispersion.stats["launches"]
    assert launches < opt.dispersion.stats["hits"]          # the cache/prefetch does its job


def test_optimize_default_config1(es):
    """BASELINE.json configs[0]: optimize_stencil.py with no flags (81 starts, N=50 -> 200)."""
    opt, x, fmin, z = run_optimize(es, "default")
    # the optimum sits on the CFL constraint: dt_ok ~ 0 within the coefficient agreement
    assert abs(np.asarray(opt.dispersion_high.dt_ok(x)).ravel()[0] - z["default_dt_ok"][0]) < 5e-4
    assert np.asarray(opt.dispersion_high.stencil_ok(x)).ravel()[0] > 0


@pytest.mark.parametrize("name", ["divfree", "divfree3d_small"])
def test_exact_fd_jacobians_leave_the_search_unchanged(es, name, capsys, monkeypatch):
    """The `jac` callables handed to SLSQP reproduce scipy's approx_derivative('2-point', abs_step=eps)
    bit for bit, so every printed line and the result are identical with and without them."""
    runs = {}
    for flag in (False, True):
        monkeypatch.setattr(es.Optimize, "exact_fd_jac", flag)
        opt = es.Optimize(default_args(**OPT_RUNS[name]))
        assert ("jac" in opt.constraints[0]) == flag
        x, fmin = opt.optimize()
        runs[flag] = (x, fmin, capsys.readouterr().out, opt.dispersion.stats["launches"])
    assert np.array_equal(runs[False][0], runs[True][0]) and runs[False][1] == runs[True][1]
    assert runs[False][2] == runs[True][2]
    assert runs[False][3] == runs[True][3]


def test_fd_jacobians_equal_scipy_approx_derivative(es):
    from scipy.optimize._numdiff import approx_derivative
    from optimize_stencil_b200.extended_stencil.dispersion import FD_STEP
    rng = np.random.default_rng(5)
    d = es.Dispersion2D(1.25, 24, es.StencilFree2D())
    pts = [np.array([0.5, 0.05, -0.02, -0.1, -0.12]), np.array([0.9, 0.3, 0.2, 0.1, 0.1]),   # usable / gated
           np.array([0.5, -0.0, 0.0, -0.1, 1e9])]                                               # -0.0 and dx == 0
    pts += [np.array([0.6, 0, 0, -0.1, -0.1]) + 0.05 * rng.standard_normal(5) for _ in range(6)]
    with np.errstate(all="ignore"):
        for p in pts:
            for fun, jac in ((d.norm, d.norm_fd), (d.stencil_ok, d.stencil_ok_fd), (d.dt_ok, d.dt_ok_fd)):
                want = approx_derivative(fun, p, method="2-point", abs_step=FD_STEP)
                np.testing.assert_array_equal(jac(p), want)
            for con in (es.make_single_max_constraint(2, 0.25), es.make_single_min_constraint(4, -1.0)):
                np.testing.assert_array_equal(con.jac(p), approx_derivative(con, p, method="2-point", abs_step=FD_STEP))


def test_stacked_constraints_are_the_reference_list_row_for_row(es):
    """optimize.py: _stacked_constraints hands SLSQP the reference's scalar constraints (optimize.py:96-110) as one
    vector; values and forward-difference normals must be those of the list, row for row, bit for bit -- also
    where SciPy falls back to a relative step (x_i + h == x_i) and on a bound."""
    from scipy.optimize._numdiff import approx_derivative
    from optimize_stencil_b200.extended_stencil.dispersion import FD_STEP
    from optimize_stencil_b200.extended_stencil.optimize import _stacked_constraints
    monkey = dict(exact_fd_jac=es.Optimize.exact_fd_jac)
    es.Optimize.exact_fd_jac = False
    try:
        opt = es.Optimize(default_args(Ngrid_low=20, Ngrid_high=20))
    finally:
        es.Optimize.exact_fd_jac = monkey["exact_fd_jac"]
    cons = opt.constraints                                    # the reference's list: no 'jac' keys
    assert len(cons) == 2 + 2 * len(opt.stencil.Parameters._fields) and "jac" not in cons[0]
    stacked = _stacked_constraints(opt.dispersion, cons)
    pts = [np.array([0.5, 0.05, -0.02, -0.1]), np.array([1.0, -1.0, 0.25, -1.0]),      # on the bounds
           np.array([0.5, 0.05, -0.02, 1e9]), np.array([0.3, 0.0, -0.0, 0.1])]          # relative-step fallback, zeros
    with np.errstate(all="ignore"):
        for p in pts:
            want = np.concatenate([np.atleast_1d(np.asarray(c["fun"](p), dtype=float)) for c in cons])
            np.testing.assert_array_equal(stacked.fun(p), want)
            J = np.vstack([np.atleast_2d(approx_derivative(c["fun"], p, method="2-point", abs_step=FD_STEP)) for c in cons])
            np.testing.assert_array_equal(stacked.jac(p), J)


def test_fd_memo_follows_the_weight(es):
    """The finite-difference points of the last iterate are memoised; a new weight must drop them."""
    d = es.Dispersion2D(1, 24, es.StencilSymmetric2D())
    x = np.array([0.5, 0.05, -0.05, -0.04])
    g1 = d.norm_fd(x)
    np.testing.assert_array_equal(d.norm_fd(x), g1)
    launches = d.stats["launches"]
    d.stencil_ok_fd(x), d.dt_ok_fd(x)
    assert d.stats["launches"] == launches                   # all three Jacobians from one launch
    d.w = 2.0
    g2 = d.norm_fd(x)
    assert d.stats["launches"] == launches + 1
    assert np.max(np.abs(g2 - 2 * g1)) <= 1e-6 * np.max(np.abs(g1))   # differences of doubled sums: same to FD noise


def test_lockstep_scan_prints_what_the_sequential_scan_prints(es, capsys):
    """All searches concurrently with one (segmented) launch per round vs one search after the other: same lines,
    same result, same evaluated points, far fewer launches.  (GPU bit-identity of the segments themselves:
    tests/test_gpu_round2.py.)"""
    runs = {}
    for flag in (False, True):
        es.Optimize.lockstep = flag
        try:
            opt = es.Optimize(default_args(div_free=True, dtrange=[0.6718, 1.0], N=4, Ngrid_low=30, Ngrid_high=60))
            x, f = opt.optimize()
        finally:
            es.Optimize.lockstep = True
        runs[flag] = (x, f, capsys.readouterr().out, dict(opt.dispersion.stats), opt.scan_stats)
    assert runs[True][2] == runs[False][2] and runs[True][2].count("x0=") == 16 + 2
    assert np.array_equal(runs[True][0], runs[False][0]) and runs[True][1] == runs[False][1]
    assert runs[True][3] == runs[False][3]
    assert runs[True][4]["launches"] * 3 < runs[True][3]["launches"] and runs[False][4] == {}


def test_lockstep_propagates_errors_from_a_search(es, monkeypatch):
    opt = es.Optimize(default_args(div_free=True, dtrange=[0.6718, 1.0], N=3, Ngrid_low=20, Ngrid_high=20))

    def boom(self, dispersion, constraints, x0):
        dispersion.norm(x0)                  # take part in a round first, then fail
        if x0[1] > -0.5:
            raise ValueError("search failed on purpose")
        return es.Optimize._minimize.__wrapped__(self, dispersion, constraints, x0)

    boom.__wrapped__ = es.Optimize._minimize
    monkeypatch.setattr(es.Optimize, "_minimize", boom)
    with pytest.raises(ValueError, match="on purpose"):
        opt.optimize()


def test_constraint_return_types_follow_the_reference(es):
    """dispersion.py:299-307 / 377-385: a negative grid minimum is returned as the NumPy scalar np.min gave, the
    corner minimum as the shape-(1,) array the record fields give; dt_ok passes a negative stencil_ok through."""
    d = es.Dispersion2D(1, 20, es.StencilFree2D())
    ok = [0.5, 0.05, -0.02, -0.1, -0.12]
    bad = [0.3, 0.3, -0.3, 0.3, 0.7]                   # tests/test_dispersion.py:39-47 (negative sqrtarg)
    assert isinstance(d.stencil_ok(ok), np.ndarray) and d.stencil_ok(ok).shape == (1,)
    assert isinstance(d.dt_ok(ok), np.ndarray) and d.dt_ok(ok).shape == (1,)
    s = d.stencil_ok(bad)
    assert isinstance(s, np.float64) and np.ndim(s) == 0 and s < 0
    t = d.dt_ok(bad)
    assert isinstance(t, np.float64) and t == s
    assert "%s" % s == "%s" % np.float64(s) and "[" not in "%s" % s     # prints like the reference's scalar


def test_fd_points_cost_one_launch_when_only_the_iterate_is_cached(es):
    """x cached without its neighbours (a start taken from a batch, prefetch off) must cost ONE launch
    of the n + 1 points, not one launch per neighbour -- and the difference never mixes two launches."""
    d = es.Dispersion2D(1, 30, es.StencilIsotropicDivFree2D())
    x = np.array([0.7, -0.02])
    d.evaluate_batch(np.array([x, x + 0.01]))          # x is cached, its neighbours are not
    launches, evaluated = d.stats["launches"], d.stats["evaluated"]
    g = d.norm_fd(x)
    assert d.stats["launches"] == launches + 1 and d.stats["evaluated"] == evaluated + 3
    d.stencil_ok_fd(x), d.dt_ok_fd(x)
    assert d.stats["launches"] == launches + 1
    fresh = es.Dispersion2D(1, 30, es.StencilIsotropicDivFree2D())
    np.testing.assert_array_equal(g, fresh.norm_fd(x))
    d2 = es.Dispersion2D(1, 30, es.StencilIsotropicDivFree2D())
    d2.fd_prefetch = False
    d2.norm(x)
    assert d2.stats["evaluated"] == 1
    np.testing.assert_array_equal(d2.norm_fd(x), g)
    assert d2.stats["launches"] == 2 and d2.stats["evaluated"] == 4


def test_views_share_the_context_and_keep_their_own_cache(es):
    d = es.Dispersion3D(1.0, 1.0, 10, es.StencilDivFree3D())
    es.weight_functions["xaxis_soft"](0.4)(d)
    v1, v2 = d.view(), d.view()
    p = [0.3, -0.10, -0.60, 0.2]
    assert v1.norm(p) == v2.norm(p) == d.norm(p)
    assert v1._context() is d._context() and v1._cache is not v2._cache
    assert v1.stats["launches"] == 1 and d.stats["launches"] == 1
    v1.parameters = p
    assert d.parameters is None or not np.array_equal(d.parameters, [9, 9, 9, 9])
    calls = []
    v3 = d.view(lambda C: calls.append(len(C)) or d._context().objective_batch(C))
    assert v3.norm(p) == d.norm(p) and calls == [5]


def test_start_grid(es):
    opt = es.Optimize(default_args(N=2, scan_dt=False))
    starts = opt.start_grid()
    assert len(starts) == 8 and starts[0][0] is None
    assert np.allclose(sorted({s[1] for s in starts}), np.linspace(-1, 1, 4)[1:-1])


def test_worker_processes_give_the_same_answer(monkeypatch, capsys):
    """The optional process pool (args.workers) only spreads SciPy's Python overhead: same starts, same
    deterministic searches, same result.  CPU variant: fork, so that the oracle-backed context is inherited."""
    fake_context.install(monkeypatch)
    monkeypatch.setenv("OPTSTENCIL_START_METHOD", "fork")
    from optimize_stencil_b200 import extended_stencil as es
    over = dict(div_free=True, dtrange=[0.6718, 1.0], Ngrid_low=30, Ngrid_high=30, N=4)
    x1, f1 = es.Optimize(default_args(**over)).optimize()
    x2, f2 = es.Optimize(default_args(workers=2, singlecore=False, **over)).optimize()
    assert f1 == f2 and np.array_equal(x1, x2)
    assert capsys.readouterr().out.count("x0=") == 2 * 16


def test_serving_order_does_not_matter(monkeypatch, capsys):
    """Workers are served as they become ready (default) or all at once per round (OPTSTENCIL_SERVE=all): segments make
    the composition of a launch irrelevant, so every printed line is the same; the stage trace goes to stderr."""
    fake_context.install(monkeypatch)
    from optimize_stencil_b200 import _capi, extended_stencil as es
    monkeypatch.setenv("OPTSTENCIL_START_METHOD", "fork")
    monkeypatch.setattr(_capi, "cuda_touched", lambda: False)     # earlier tests of this process asked for the device count
    monkeypatch.setenv("OPTSTENCIL_TRACE", "1")
    over = dict(div_free=True, dtrange=[0.6718, 1.0], Ngrid_low=24, Ngrid_high=24, N=4, workers=3, singlecore=False)
    outs, stats = [], []
    for at_once in (False, True):
        monkeypatch.setattr(es.Optimize, "serve_all_at_once", at_once)
        opt = es.Optimize(default_args(**over))
        x, f = opt.optimize()
        cap = capsys.readouterr()
        outs.append((cap.out, tuple(x), f))
        stats.append(opt.scan_stats)
        assert "workers served" in cap.err and "scan done" in cap.err
    assert outs[0] == outs[1]
    assert stats[0]["rows"] == stats[1]["rows"] and stats[0]["start_method"] == "fork"
    assert stats[0]["launches"] >= stats[1]["launches"]          # at most as much merging as with a barrier per round
    # a process that has called into CUDA is not forked, whatever the environment says: the workers are spawned
    monkeypatch.setattr(_capi, "cuda_touched", lambda: True)
    opt = es.Optimize(default_args(**over))
    x, f = opt.optimize()
    cap = capsys.readouterr()
    assert opt.scan_stats["start_method"] == "spawn" and "spawning the workers instead" in cap.out
    assert (cap.out.replace("optimize: not forking a process that has called into CUDA; spawning the workers instead\n", ""),
            tuple(x), f) == outs[0]


def test_fork_only_before_the_first_cuda_call():
    """`_capi.cuda_touched()` is what `Optimize._start_workers` goes by: False in a fresh process, True after the first
    call that reaches the CUDA runtime (here: the device count, which works without a GPU)."""
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from optimize_stencil_b200 import _capi\n"
            "assert not _capi.cuda_touched()\n"
            "_capi.device_count()\n"
            "assert _capi.cuda_touched()\n"
            "print('ok')\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr


@pytest.mark.gpu
def test_worker_processes_on_the_gpu(capsys):
    """Same with real spawned workers (SciPy only) served by this process, which owns the device."""
    from optimize_stencil_b200 import extended_stencil as es
    over = dict(div_free=True, dtrange=[0.6718, 1.0], Ngrid_low=30, Ngrid_high=30, N=4)
    x1, f1 = es.Optimize(default_args(**over)).optimize()
    x2, f2 = es.Optimize(default_args(workers=2, singlecore=False, **over)).optimize()
    assert f1 == f2 and np.array_equal(x1, x2)
