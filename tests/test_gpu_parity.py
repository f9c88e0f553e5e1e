"""GPU parity of the CUDA path (through the C-ABI) against the oracle and the golden fixtures.

Bars (BASELINE.json north_star / SURVEY.md 8d):
  * objective  |S_gpu - S_ref| / |S_ref| <= 1e-12
  * omega(k)   max relative error <= 1e-13 (k = 0 excluded where omega = 0: absolute there)
  * min / max of sqrtarg: BITWISE (they steer SLSQP through stencil_ok / dt_ok)
Device contract for the extrema: Amin = min(sqrtarg U {+0}), Amax = max(sqrtarg U {+0}); sqrtarg is
exactly 0 at k = 0, so for any x-range containing plane 0 these are numpy's min / max up to the sign
of zero.
"""
import os

import numpy as np
import pytest

import np_oracle as O
from conftest import relerr

pytestmark = pytest.mark.gpu

TOL_S = 1e-12
TOL_OMEGA = 1e-13


@pytest.fixture(scope="module")
def sb():
    import optimize_stencil_b200 as m
    assert m.device_count() >= 1, "no CUDA device visible"
    return m


def ctx_from_fixture(sb, c):
    g = c["grid"]
    dim = c["dim"]
    return sb.Context(dim, [g["cos"][a] for a in range(dim)], [g["s2"][a] for a in range(dim)],
                      [g["kcomp"][a] for a in range(dim)], Y=c["Y"], Z=c["Z"])


def ctx_from_grid(sb, grid):
    cos, s2, kc = grid.flat_tables()
    return sb.Context(grid.dim, cos, s2, kc, Y=grid.Y, Z=grid.Z)


def clamp_min(a):
    return a if a < 0 else 0.0


def same_bits(a, b):
    a, b = np.float64(a), np.float64(b)
    return (a == b) or (np.isnan(a) and np.isnan(b))


def test_golden_cases_objective_and_extrema(sb, golden):
    """Every fixture case produced by the unmodified reference, fed with the reference's own tables."""
    checked = 0
    for i in range(golden.n):
        c = golden.case(i)
        if np.any(np.isnan(c["params"])):
            continue  # NaN parameters never reach the device (dispersion.py:241-243)
        tag = "case %d %s" % (i, c["tag"])
        with ctx_from_fixture(sb, c) as ctx:
            grid = O.Grid(c["dim"], c["N"], c["Y"], c["Z"])
            w = O.weight(grid, c["wkind"], c["wparams"])
            if c["wkind"] != "equal":
                ctx.set_weight(w)
            S, Amin, Amax, nan = ctx.objective_batch(c["coeffs"])
            assert same_bits(Amin[0], clamp_min(c["Amin"])), (tag, Amin[0], c["Amin"])
            assert same_bits(Amax[0], max(c["Amax"], 0.0)), (tag, Amax[0], c["Amax"])
            if not c["gated"]:
                assert nan[0] == 0, tag
                assert relerr(S[0], c["S"]) <= TOL_S, (tag, S[0], c["S"])
                om, amin2, amax2, nan2 = ctx.export(c["coeffs"], 0)
                assert same_bits(amin2, Amin[0]) and same_bits(amax2, Amax[0]) and nan2 == 0, tag
                idx = golden.sample_indices(i, om.size)
                err = relerr(om.ravel()[idx], c["omega_samples"])
                assert np.max(err) <= TOL_OMEGA, (tag, np.max(err))
                if c["omega_full"] is not None:
                    assert om.shape == c["omega_full"].shape
                    assert np.max(relerr(om, c["omega_full"])) <= TOL_OMEGA, tag
                checked += 1
    assert checked >= 50


def test_live_oracle_bitwise_sqrtarg_and_sqrtres(sb):
    """sqrtarg is evaluated in the reference's operation order: dense export equals NumPy bit for bit."""
    rng = np.random.default_rng(3)
    for dim, N, Y, Z, fam in [(3, 33, 1.0, 1.0, "StencilFree3D"), (3, 20, 1.3, 0.7, "StencilFree3D"),
                              (2, 77, 1.0, 1.0, "StencilFree2D"), (2, 50, 2.0, 1.0, "StencilFree2D")]:
        grid = O.Grid(dim, N, Y, Z)
        with ctx_from_grid(sb, grid) as ctx:
            for _ in range(3):
                p = 0.1 * rng.standard_normal(len(O.parameter_names(fam)))
                p[0] = 0.35
                co = O.coefficients(fam, p)
                A_ref = O.sqrtarg(grid, co)
                A, amin, amax, _ = ctx.export(co, 1)
                np.testing.assert_array_equal(A, A_ref)
                assert same_bits(amin, clamp_min(A_ref.min())) and same_bits(amax, max(A_ref.max(), 0.0))
                with np.errstate(invalid="ignore"):
                    R_ref = np.sqrt(A_ref)
                R, _, _, nn = ctx.export(co, 2)
                np.testing.assert_array_equal(R, R_ref)
                assert nn == np.count_nonzero(np.isnan(R_ref))


def test_batch_matches_singles_and_oracle(sb):
    """A mixed batch (feasible, infeasible, on-CFL) against the oracle's un-gated device contract."""
    rng = np.random.default_rng(11)
    grid = O.Grid(3, 40)
    xc = np.array([0.5, .02, .03, .01, .02, .03, .01, -.02, -.03, -.01])
    P = [xc + 1e-3 * rng.standard_normal(10) for _ in range(13)]
    P.append(np.array([0.4, .6, -.5, .4, .3, -.7, .2, .24, -.9, .1]))   # negative sqrtarg
    P.append(np.array([1.9, 0, 0, 0, 0, 0, 0, 0, 0, 0]))               # dt beyond CFL: s > 1 -> NaNs
    co = np.stack([O.coefficients("StencilFree3D", p) for p in P])
    with ctx_from_grid(sb, grid) as ctx:
        S, Amin, Amax, nan = ctx.objective_batch(co)
        S2, Amin2, Amax2, nan2 = ctx.objective_batch(co)
        for a, b in ((S, S2), (Amin, Amin2), (Amax, Amax2)):
            np.testing.assert_array_equal(a, b)  # deterministic
        np.testing.assert_array_equal(nan, nan2)
        for b in range(len(P)):
            Sr, Ar_min, Ar_max, nr = O.device_contract(grid, co[b])
            assert same_bits(Amin[b], clamp_min(Ar_min)) and same_bits(Amax[b], max(Ar_max, 0.0)), b
            assert (nan[b] > 0) == (nr > 0), (b, nan[b], nr)
            if nr == 0:
                assert relerr(S[b], Sr) <= TOL_S, (b, S[b], Sr)
            else:
                assert np.isnan(S[b])
            s1 = ctx.objective_batch(co[b])
            assert same_bits(s1[1][0], Amin[b]) and same_bits(s1[2][0], Amax[b]) and (s1[3][0] > 0) == (nan[b] > 0)
            if nr == 0:
                assert relerr(s1[0][0], S[b]) <= 1e-13


@pytest.mark.parametrize("cand,ty,tx", [(1, 8, 1), (2, 16, 3), (4, 64, 7), (8, 32, 2), (4, 256, 64)])
def test_tilings_agree(sb, cand, ty, tx):
    """Tile shape only changes the (fixed) summation order: extrema bitwise, S to 1e-13."""
    rng = np.random.default_rng(5)
    grid = O.Grid(3, 36, 1.0, 1.0)
    xc = np.array([0.5, .02, .03, .01, .02, .03, .01, -.02, -.03, -.01])
    co = np.stack([O.coefficients("StencilFree3D", xc + 1e-3 * rng.standard_normal(10)) for _ in range(7)])
    with ctx_from_grid(sb, grid) as ctx:
        ref = ctx.objective_batch(co)
        ctx.set_tiling(cand, ty, tx)
        got = ctx.objective_batch(co)
        np.testing.assert_array_equal(got[1], ref[1])
        np.testing.assert_array_equal(got[2], ref[2])
        np.testing.assert_array_equal(got[3] > 0, ref[3] > 0)
        assert np.max(relerr(got[0], ref[0])) <= 1e-13


def test_weights_plane_and_dense(sb):
    """weight.py:4-31 shapes: (1,N,N) plane, N^3 dense, (1,N) and (N,N) in 2D."""
    for dim, N in ((3, 24), (2, 50)):
        grid = O.Grid(dim, N, 1.2, 0.9)
        fam = "StencilFree3D" if dim == 3 else "StencilFree2D"
        p = np.zeros(len(O.parameter_names(fam)))
        p[0] = 0.3
        p[-1] = -0.05
        co = O.coefficients(fam, p)
        with ctx_from_grid(sb, grid) as ctx:
            for kind, params in (("xaxis", ()), ("xaxis_soft", ()), ("xaxis_soft", (0.5,)), ("equal", ())):
                w = O.weight(grid, kind, params)
                if kind == "equal":
                    ctx.set_weight_unit()
                else:
                    ctx.set_weight(w)
                S = ctx.objective_batch(co)[0][0]
                Sr = O.device_contract(grid, co, w)[0]
                assert relerr(S, Sr) <= TOL_S, (dim, kind, S, Sr)
            rng = np.random.default_rng(2)
            w = rng.random(grid.shape)
            ctx.set_weight(w)
            assert relerr(ctx.objective_batch(co)[0][0], O.device_contract(grid, co, w)[0]) <= TOL_S


def test_x_slabs_partition_the_grid(sb):
    grid = O.Grid(3, 48)
    co = O.coefficients("StencilFree3D", [0.5, .02, .03, .01, .02, .03, .01, -.02, -.03, -.01])
    with ctx_from_grid(sb, grid) as ctx:
        full = ctx.objective_batch(co)
        cuts = [0, 5, 6, 30, 48]
        parts = [ctx.objective_batch(co, a, b) for a, b in zip(cuts[:-1], cuts[1:])]
        assert relerr(sum(p[0][0] for p in parts), full[0][0]) <= 1e-13
        assert same_bits(min(p[1][0] for p in parts), full[1][0])
        assert same_bits(max(p[2][0] for p in parts), full[2][0])
        for (a, b), p in zip(zip(cuts[:-1], cuts[1:]), parts):  # each slab against the slab-wise oracle
            Sr, amin, amax, _ = O.device_contract(O.Grid(3, 48, xslab=(a, b)), co)
            assert relerr(p[0][0], Sr) <= TOL_S and same_bits(p[2][0], max(amax, 0.0))
        om = np.concatenate([ctx.export(co, 0, a, b)[0] for a, b in zip(cuts[:-1], cuts[1:])])
        np.testing.assert_array_equal(om, ctx.export(co, 0)[0])


def test_config2_lehe_256(sb, golden):
    """BASELINE.json configs[1]: calculate_omega.py --lehe --params 0.96 on a 256^3 grid."""
    c = [golden.case(i) for i in range(golden.n) if golden.case(i)["tag"] == "lehe-3d-256"][0]
    grid = O.Grid(3, 256)
    with ctx_from_grid(sb, grid) as ctx:  # tables from this machine's NumPy, as the product does
        S, Amin, Amax, nan = ctx.objective_batch(c["coeffs"])
        assert relerr(S[0], c["S"]) <= TOL_S and nan[0] == 0
        assert relerr(S[0] * O.norm_scale(3, 256), c["norm"]) <= TOL_S
        om = ctx.export(c["coeffs"], 0)[0]
        idx = golden.sample_indices(c["id"], om.size)
        assert np.max(relerr(om.ravel()[idx], c["omega_samples"])) <= TOL_OMEGA
        ref = O.evaluate(grid, c["coeffs"])  # the full 16.8 M-point field against the live oracle
        assert np.max(relerr(om, ref["omega"])) <= TOL_OMEGA
        assert relerr(S[0], ref["S"]) <= TOL_S


def test_full_size_512_properties(sb):
    """BASELINE.json configs[3] grid size (512^3): size-independent properties + a slab-wise oracle."""
    N = 512
    grid_tabs = O.Grid(3, 8)  # only to reuse the table code path below
    x = np.linspace(0, np.pi, N)
    cos = [np.cos(x)] * 3
    s2 = [0.5 * (1. - np.cos(x))] * 3
    rng = np.random.default_rng(0)
    xc = np.array([0.5, .02, .03, .01, .02, .03, .01, -.02, -.03, -.01])
    P = [xc + 1e-3 * rng.standard_normal(10) for _ in range(8)]
    co = np.stack([O.coefficients("StencilFree3D", p) for p in P])
    with sb.Context(3, cos, s2, [x, x, x]) as ctx:
        full = ctx.objective_batch(co)
        assert np.all(full[3] == 0) and np.all(np.isfinite(full[0]))
        # (1) slabs partition the sum and the extrema
        halves = [ctx.objective_batch(co, 0, 200), ctx.objective_batch(co, 200, 512)]
        assert np.max(relerr(halves[0][0] + halves[1][0], full[0])) <= 1e-13
        np.testing.assert_array_equal(np.maximum(halves[0][2], halves[1][2]), full[2])
        # (2) a permutation of the axes of a fully anisotropic stencil leaves the objective unchanged
        #     (x<->y: alphas, deltas and the beta index pairs swap roles)
        def swap_xy(c):
            dt, ax, ay, az, bxy, bxz, byx, byz, bzx, bzy, dx_, dy_, dz_ = c
            return np.array([dt, ay, ax, az, byx, byz, bxy, bxz, bzy, bzx, dy_, dx_, dz_])
        sw = ctx.objective_batch(np.stack([swap_xy(c) for c in co]))
        assert np.max(relerr(sw[0], full[0])) <= 1e-12
        assert np.max(relerr(sw[2], full[2])) <= 1e-15
        # (3) slab-wise NumPy oracle on three thin slabs of the full-size grid
        for a, b in ((0, 4), (255, 258), (508, 512)):
            got = ctx.objective_batch(co[:2], a, b)
            for j in range(2):
                Sr, amin, amax, nn = O.device_contract(O.Grid(3, N, xslab=(a, b)), co[j])
                assert relerr(got[0][j], Sr) <= TOL_S and same_bits(got[2][j], max(amax, 0.0)) and nn == 0


def test_argument_errors(sb):
    grid = O.Grid(2, 16)
    with ctx_from_grid(sb, grid) as ctx:
        co = O.coefficients("StencilFree2D", [0.3, 0, 0, 0, 0])
        with pytest.raises(sb.StencilB200Error):
            ctx.objective_batch(co, 0, 8)       # x-slabs are a 3-D feature
        with pytest.raises(ValueError):
            ctx.objective_batch(np.zeros((2, 13)))
        with pytest.raises(sb.StencilB200Error):
            ctx.set_tiling(5, 0, 0)
    with pytest.raises(sb.StencilB200Error):
        sb.Context(3, [np.ones(4)] * 3, [np.zeros(4)] * 3, [np.zeros(4)] * 3, device=99)


@pytest.mark.parametrize("dim,N", [(3, 1), (3, 2), (3, 5), (3, 31), (3, 33), (3, 65), (2, 1), (2, 2), (2, 3), (2, 257), (2, 300)])
def test_ragged_and_tiny_grids(sb, dim, N):
    """Grid sizes that do not fill warps, 32-row tiles or z tiles; N = 1 is the origin alone."""
    fam = "StencilFree3D" if dim == 3 else "StencilFree2D"
    rng = np.random.default_rng(N)
    P = 0.05 * rng.standard_normal((3, len(O.parameter_names(fam))))
    P[:, 0] = [0.2, 0.45, 0.6]
    if N == 1:
        x = np.array([0.0])
        tabs = ([np.cos(x)] * dim, [0.5 * (1. - np.cos(x))] * dim, [x] * dim)
        grid = O.Grid.from_tables(dim, *tabs)
    else:
        grid = O.Grid(dim, N)
        tabs = grid.flat_tables()
    co = np.stack([O.coefficients(fam, p) for p in P])
    with sb.Context(dim, *tabs) as ctx:
        S, Amin, Amax, nan = ctx.objective_batch(co)
        for b in range(3):
            Sr, amin, amax, nr = O.device_contract(grid, co[b])
            assert same_bits(Amin[b], clamp_min(amin)) and same_bits(Amax[b], max(amax, 0.0)), (N, b)
            assert (nan[b] > 0) == (nr > 0)
            if nr == 0:
                assert (S[b] == Sr == 0.0) or relerr(S[b], Sr) <= TOL_S, (N, b, S[b], Sr)
        om = ctx.export(co[0], 0)[0]
        with np.errstate(invalid="ignore"):
            ref = O.omega_from_sqrtarg(np.broadcast_to(O.sqrtarg(grid, co[0]), grid.shape), co[0][0])
        assert np.max(relerr(om, ref)) <= TOL_OMEGA


@pytest.mark.parametrize("B", [1, 2, 3, 4, 5, 7, 8, 9, 63, 64, 65, 130])
def test_batch_sizes_and_padding(sb, B):
    """Batches that do not fill a candidate chunk; both the zero-copy (B <= 64) and the staged path."""
    rng = np.random.default_rng(B)
    grid = O.Grid(3, 20)
    xc = np.array([0.5, .02, .03, .01, .02, .03, .01, -.02, -.03, -.01])
    co = np.stack([O.coefficients("StencilFree3D", xc + 2e-3 * rng.standard_normal(10)) for _ in range(B)])
    with ctx_from_grid(sb, grid) as ctx:
        S, Amin, Amax, nan = ctx.objective_batch(co)
        pick = sorted(set([0, B // 2, B - 1]))
        for b in pick:
            Sr, amin, amax, nr = O.device_contract(grid, co[b])
            assert relerr(S[b], Sr) <= TOL_S and same_bits(Amax[b], amax) and nan[b] == 0
        one = np.array([ctx.objective_batch(co[b])[0][0] for b in pick])
        assert np.max(relerr(one, S[pick])) <= 1e-13


def test_nan_and_inf_coefficients_do_not_poison_neighbours(sb):
    """A NaN / inf coefficient row yields NaN for that row only (the host gate never sends NaN parameters,
    but the C-ABI must not trap on them)."""
    grid = O.Grid(3, 18)
    good = O.coefficients("StencilFree3D", [0.5, .02, .03, .01, .02, .03, .01, -.02, -.03, -.01])
    bad1, bad2 = good.copy(), good.copy()
    bad1[5] = np.nan
    bad2[10] = np.inf
    co = np.stack([good, bad1, good, bad2, good])
    with ctx_from_grid(sb, grid) as ctx:
        S, Amin, Amax, nan = ctx.objective_batch(co)
        # rows 0 and 2 share a candidate chunk with the NaN row (whose CTA evaluates sqrtarg in the reference
        # order everywhere), row 4 does not: same value up to the rounding of the sum, same extrema bit for bit
        assert S[0] == S[2] and relerr(S[4], S[0]) <= 1e-13 and np.isfinite(S[0]) and nan[0] == 0
        assert same_bits(Amax[0], Amax[4]) and same_bits(Amin[0], Amin[4]) and same_bits(Amax[0], Amax[2])
        assert np.isnan(S[1]) and np.isnan(Amin[1]) and np.isnan(Amax[1]) and nan[1] > 0
        assert np.isnan(S[3]) and nan[3] > 0
        assert relerr(S[0], O.device_contract(grid, good)[0]) <= TOL_S


def test_config3_scan_grid(sb):
    """BASELINE.json configs[2] in miniature: the --div-free 3-D family (dt, deltax), a dt x deltax start
    grid in ONE launch, including gated candidates, against the oracle looped over the same points."""
    grid = O.Grid(3, 40)
    dts = np.linspace(0.40, 0.80, 10)[1:-1]     # straddles the 3-D CFL limit of this family
    dls = np.linspace(-0.3, 0.25, 10)[1:-1]
    P = np.array([(a, b) for a in dts for b in dls])
    co = np.stack([O.coefficients("StencilIsotropicDivFree3D", p) for p in P])
    with ctx_from_grid(sb, grid) as ctx:
        S, Amin, Amax, nan = ctx.objective_batch(co)
    gated = 0
    for b in range(len(P)):
        r = O.evaluate(grid, co[b])
        assert same_bits(Amin[b], clamp_min(r["Amin"])) and same_bits(Amax[b], max(r["Amax"], 0.0)), b
        if r["omega"] is None:
            gated += 1
        else:
            assert relerr(S[b], r["S"]) <= TOL_S and nan[b] == 0, b
    assert 0 < gated < len(P)


def test_config5_grid_2048_slabs(sb):
    """BASELINE.json configs[4] grid size (2048^3, 8.6e9 k-points, not representable in the reference):
    x-slabs partition the sum and the extrema; thin slabs agree with the slab-wise oracle."""
    N = 2048
    x = np.linspace(0, np.pi, N)
    cos, s2 = np.cos(x), 0.5 * (1. - np.cos(x))
    co = O.coefficients("StencilFree3D", [0.5, .02, .03, .01, .02, .03, .01, -.02, -.03, -.01])
    with sb.Context(3, [cos] * 3, [s2] * 3, [x] * 3) as ctx:
        full = ctx.objective_batch(co)
        assert full[3][0] == 0 and np.isfinite(full[0][0])
        cuts = [0, 256, 1024, 1027, 2048]          # the 8-GPU slab size, a big one, a ragged one
        parts = [ctx.objective_batch(co, a, b) for a, b in zip(cuts[:-1], cuts[1:])]
        assert relerr(sum(p[0][0] for p in parts), full[0][0]) <= 1e-13
        assert same_bits(max(p[2][0] for p in parts), full[2][0])
        for a, b in ((0, 1), (1024, 1026), (2047, 2048)):
            got = ctx.objective_batch(co, a, b)
            Sr, amin, amax, nn = O.device_contract(O.Grid(3, N, xslab=(a, b)), co)
            assert relerr(got[0][0], Sr) <= TOL_S and same_bits(got[2][0], max(amax, 0.0)) and nn == 0


def test_full_size_512_against_the_c_oracle(sb):
    """BASELINE.json configs[3] at its full grid size: whole-grid sums and extrema of several candidates
    against the plain-C oracle (compensated summation, libm sqrt/asin) -- no sampling, all 134 M k-points."""
    import c_oracle
    N = 512
    x = np.linspace(0, np.pi, N)
    cos, s2 = np.cos(x), 0.5 * (1. - np.cos(x))
    rng = np.random.default_rng(0)
    xc = np.array([0.5, .02, .03, .01, .02, .03, .01, -.02, -.03, -.01])
    P = xc + 1e-3 * rng.standard_normal((6, 10))        # the first candidates of bench.py's batch
    co = np.stack([O.coefficients("StencilFree3D", p) for p in P])
    ref = c_oracle.COracle(3, [cos] * 3, [s2] * 3, [x] * 3).objective(co[:3])
    with sb.Context(3, [cos] * 3, [s2] * 3, [x] * 3) as ctx:
        S, Amin, Amax, nan = ctx.objective_batch(co)
        assert np.all(nan == 0)
        for b in range(3):
            assert relerr(S[b], ref[0][b]) <= TOL_S, (b, S[b], ref[0][b])
            assert same_bits(Amax[b], ref[2][b]) and same_bits(Amin[b], clamp_min(ref[1][b]))


def test_non_unit_cells_divide_exactly(sb):
    """sqrtarg divides by (Y dx)^2 and (Z dx)^2 (dispersion.py:398).  The kernels use a 5-instruction
    invariant-divisor division; it must round exactly like numpy's true division, also for awkward
    divisors (mantissa of all ones, powers of two, tiny, huge)."""
    rng = np.random.default_rng(42)
    cells = [(1.25, 0.8), (np.sqrt(2 - 2 ** -52), 3.0), (0.5, 4.0), (1 / 3, 7 / 3), (1e-3, 1e3),
             (np.nextafter(1.0, 2.0), np.nextafter(1.0, 0.0)), (np.pi, np.e)]
    cells += [tuple(np.exp(rng.uniform(-2, 2, 2))) for _ in range(8)]
    for Y, Z in cells:
        grid = O.Grid(3, 24, Y, Z)
        with ctx_from_grid(sb, grid) as ctx:
            for _ in range(2):
                p = 0.1 * rng.standard_normal(10)
                p[0] = 0.3
                co = O.coefficients("StencilFree3D", p)
                A_ref = O.sqrtarg(grid, co)
                A = ctx.export(co, 1)[0]
                np.testing.assert_array_equal(A, A_ref)
                r = ctx.objective_batch(co)
                assert same_bits(r[1][0], clamp_min(A_ref.min())) and same_bits(r[2][0], max(A_ref.max(), 0.0)), (Y, Z)
        grid2 = O.Grid(2, 40, Y)
        with ctx_from_grid(sb, grid2) as ctx:
            co = O.coefficients("StencilFree2D", [0.3, 0.05, -0.04, -0.02, 0.03])
            np.testing.assert_array_equal(ctx.export(co, 1)[0], O.sqrtarg(grid2, co))


def test_fuzz_against_the_c_oracle(sb):
    """Differential fuzzing: random grids, cells, families, weights and parameter vectors (many of them
    infeasible: negative sqrtarg, s > 1, tiny and huge dt) against the plain-C oracle -- every class of
    the omega evaluation, every reduction path, the gate inputs bitwise."""
    import c_oracle
    rng = np.random.default_rng(20260101)
    fams = sorted(O.STENCILS)
    n_omega = 0
    for it in range(120):
        fam = fams[it % len(fams)]
        dim = O.STENCILS[fam][0]
        N = int(rng.integers(2, 41 if dim == 3 else 160))
        Y, Z = (1.0, 1.0) if it % 3 else (float(np.exp(rng.uniform(-1, 1))), float(np.exp(rng.uniform(-1, 1))))
        grid = O.Grid(dim, N, Y, Z)
        cos, s2, kc = grid.flat_tables()
        dof = len(O.parameter_names(fam))
        B = int(rng.integers(1, 10))
        P = rng.uniform(-0.4, 0.4, (B, dof)) * rng.choice([0.05, 0.3, 1.0], (B, 1))
        P[:, 0] = rng.choice([0.01, 0.2, 0.5, 0.9, 1.3], B) * rng.uniform(0.8, 1.2, B)
        co = np.stack([O.coefficients(fam, p) for p in P])
        w = None
        if it % 4 == 1:
            w = rng.random((1,) + grid.shape[1:]) if dim == 3 else rng.random(grid.shape)
        elif it % 4 == 2:
            w = rng.random(grid.shape)
        ref = c_oracle.COracle(dim, cos, s2, kc, Y, Z, w)
        Sr, lo, hi, nr = ref.objective(co)
        with sb.Context(dim, cos, s2, kc, Y=Y, Z=Z) as ctx:
            if w is not None:
                ctx.set_weight(w)
            S, Amin, Amax, nan = ctx.objective_batch(co)
            for b in range(B):
                tag = (it, fam, N, Y, Z, b, P[b])
                assert same_bits(Amin[b], clamp_min(lo[b])) and same_bits(Amax[b], max(hi[b], 0.0)), tag
                assert (nan[b] > 0) == (nr[b] > 0), tag
                if nr[b] == 0:
                    assert (S[b] == 0.0 == Sr[b]) or relerr(S[b], Sr[b]) <= TOL_S, tag + (S[b], Sr[b])
                else:
                    assert np.isnan(S[b]), tag
            if it % 5 == 0:
                om = ctx.export(co[0], 0)[0]
                want = ref.export(co[0], 0)
                both_nan = np.isnan(om) & np.isnan(want)
                assert np.array_equal(np.isnan(om), np.isnan(want)), (it, fam)
                assert np.max(np.where(both_nan, 0.0, relerr(om, want))) <= TOL_OMEGA, (it, fam)
                n_omega += 1
    assert n_omega >= 20


def test_guarded_regrouping_of_sqrtarg_on_hostile_candidates(sb):
    """The objective kernel evaluates sqrtarg regrouped (3 instructions) and falls back to the reference's
    operation order where the extrema, the sign or the CFL class could depend on the last bits -- and for
    whole candidates whose coefficients defeat its error bound (csrc/objective_fa.cuh).  Raw coefficient rows
    that aim at every guard: cancellation, plateaus, zero and negative corners, huge and tiny dt."""
    rng = np.random.default_rng(77)
    yee = np.array([0.5, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0], dtype=float)
    rows = [yee]
    r = yee.copy(); r[1:4] = 0.0; rows.append(r)                                  # sqrtarg == 0 everywhere
    r = yee.copy(); r[1], r[3] = 0.0, 0.0; rows.append(r)                         # plateau: depends on y only
    r = yee.copy(); r[1:4] = [1e9, -1e9, 3.0]; r[4:10] = 1e9 * rng.standard_normal(6); rows.append(r)   # cancellation
    r = yee.copy(); r[1:4] = [-1.0, 0.5, 0.25]; rows.append(r)                    # negative towards the x corner
    r = yee.copy(); r[10:13] = [-0.4, -0.4, -0.4]; r[1:4] += 1.2; rows.append(r)  # corner near zero: 1 + 1.2 - 3*0.4*... 
    r = yee.copy(); r[0] = 1e-8; rows.append(r)                                   # tiny dt: everything in the LOW class
    r = yee.copy(); r[0] = 1e4; rows.append(r)                                    # huge dt: s > 1 almost everywhere
    r = yee.copy(); r[0] = 1.0 / np.sqrt(3.0); rows.append(r)                     # Yee ON its CFL limit (s == 1 at the corner)
    r = yee.copy(); r[0] = np.nextafter(1.0 / np.sqrt(3.0), 1.0); rows.append(r)  # one ulp beyond
    for _ in range(6):                                                            # wild random rows
        r = rng.standard_normal(13) * rng.choice([1e-3, 1.0, 1e3]); r[0] = abs(r[0]) + 0.1; rows.append(r)
    rows += [yee * 1.0, yee * 1.0]
    co = np.stack(rows)
    for (N, Y, Z, xr) in [(24, 1.0, 1.0, None), (33, 1.3, 0.7, None), (40, 1.0, 1.0, (0, 3)), (40, 1.0, 1.0, (17, 29))]:
        grid = O.Grid(3, N, Y, Z)
        with ctx_from_grid(sb, grid) as ctx:
            S, Amin, Amax, nan = ctx.objective_batch(co) if xr is None else ctx.objective_batch(co, *xr)
        sub = grid if xr is None else O.Grid.from_tables(
            3, [grid.cos[0].ravel()[xr[0]:xr[1]], grid.cos[1].ravel(), grid.cos[2].ravel()],
            [grid.s2[0].ravel()[xr[0]:xr[1]], grid.s2[1].ravel(), grid.s2[2].ravel()],
            [grid.kcomp[0].ravel()[xr[0]:xr[1]], grid.kcomp[1].ravel(), grid.kcomp[2].ravel()], Y, Z)
        for b in range(len(co)):
            Sr, amin, amax, nr = O.device_contract(sub, co[b])
            tag = (N, Y, Z, xr, b)
            assert same_bits(Amin[b], clamp_min(amin)) and same_bits(Amax[b], max(amax, 0.0)), tag + (Amin[b], amin, Amax[b], amax)
            assert (nan[b] > 0) == (nr > 0), tag
            if nr == 0:
                assert (S[b] == 0.0 == Sr) or relerr(S[b], Sr) <= TOL_S, tag + (S[b], Sr)


def test_tables_that_are_not_cosines(sb):
    """The C-ABI takes the per-axis tables from the caller.  Tables outside [-1, 1] / [0, 1] widen the error
    bound of the regrouped sqrtarg (GridDev::tab_scale); the extrema must stay the oracle's bits."""
    rng = np.random.default_rng(5)
    N = 21
    base = O.Grid(3, N)
    cos = [3.0 * c.ravel() + 0.1 * rng.standard_normal(N) for c in base.cos]
    s2 = [2.5 * t.ravel() * (1 + 0.2 * rng.random(N)) for t in base.s2]
    kc = [k.ravel().copy() for k in base.kcomp]
    grid = O.Grid.from_tables(3, cos, s2, kc)
    xc = np.array([0.2, .02, .03, .01, .02, .03, .01, -.02, -.03, -.01])
    co = np.stack([O.coefficients("StencilFree3D", xc + 0.05 * rng.standard_normal(10)) for _ in range(9)])
    with sb.Context(3, cos, s2, kc) as ctx:
        S, Amin, Amax, nan = ctx.objective_batch(co)
        np.testing.assert_array_equal(ctx.export(co[0], 1)[0], O.sqrtarg(grid, co[0]))
    for b in range(len(co)):
        Sr, amin, amax, nr = O.device_contract(grid, co[b])
        assert same_bits(Amin[b], clamp_min(amin)) and same_bits(Amax[b], max(amax, 0.0)), b
        assert (nan[b] > 0) == (nr > 0), b
        if nr == 0:
            assert relerr(S[b], Sr) <= TOL_S, (b, S[b], Sr)


_VARIANT_SCRIPT = """
import sys, numpy as np
sys.path.insert(0, %(root)r)
import optimize_stencil_b200 as sb
z = np.load(%(inp)r)
out = {}
for i in range(int(z['ncase'])):
    cos, s2, kc = [[z['c%%d_%%s%%d' %% (i, n, a)] for a in range(3)] for n in ('cos', 's2', 'kc')]
    with sb.Context(3, cos, s2, kc, Y=float(z['c%%d_Y' %% i]), Z=float(z['c%%d_Z' %% i])) as ctx:
        if 'c%%d_w' %% i in z.files:
            ctx.set_weight(z['c%%d_w' %% i])
        xr = z['c%%d_xr' %% i]
        S, lo, hi, nan = ctx.objective_batch(z['c%%d_co' %% i], int(xr[0]), int(xr[1]))
        out['S%%d' %% i], out['lo%%d' %% i], out['hi%%d' %% i], out['nan%%d' %% i] = S, lo, hi, nan
np.savez(%(outp)r, **out)
"""


def test_regrouped_kernel_agrees_with_the_reference_order_kernel(sb, tmp_path):
    """Differential test of csrc/objective_fa.cuh against the SAME library built with -DSB200_FAST_A=0 (sqrtarg in
    the reference's operation order at every k-point, libstencil_b200_reforder.so, loaded in a subprocess through
    SB200_LIB): extrema and NaN flags bit for bit, sums to 1e-13, on feasible, on-CFL, infeasible and hostile
    candidates, unit and non-unit cells, weights and x-slabs."""
    import subprocess
    import sys
    from optimize_stencil_b200 import _build
    assert os.path.exists(_build.LIB_REFORDER), "run __graft_entry__.build() first: " + _build.LIB_REFORDER
    rng = np.random.default_rng(31)
    xc = np.array([0.5, .02, .03, .01, .02, .03, .01, -.02, -.03, -.01])
    dlx = 0.25 * (1 - 1 / 0.96 ** 2 * np.sin(np.pi * 0.96 / 2.0))
    lehe = np.array([0.96, .125, .125, .125, .125, .125, .125, dlx, 0.0, 0.0])
    P = np.concatenate([xc + 1e-3 * rng.standard_normal((40, 10)), lehe + 1e-5 * rng.standard_normal((24, 10)),
                        xc + 0.3 * rng.standard_normal((40, 10)), 3.0 * rng.standard_normal((16, 10))])
    P[64:, 0] = np.abs(P[64:, 0]) + 0.05
    co = np.stack([O.coefficients("StencilFree3D", p) for p in P])
    cases = [(64, 1.0, 1.0, None, (0, 64)), (50, 1.3, 0.8, None, (0, 50)), (48, 1.0, 1.0, "plane", (0, 48)),
             (40, 0.9, 1.0, "dense", (0, 40)), (64, 1.0, 1.0, None, (5, 23)), (33, 1.0, 1.0, None, (30, 33))]
    inp = dict(ncase=len(cases))
    for i, (N, Y, Z, wk, xr) in enumerate(cases):
        grid = O.Grid(3, N, Y, Z)
        cos, s2, kc = grid.flat_tables()
        for a in range(3):
            inp["c%d_cos%d" % (i, a)], inp["c%d_s2%d" % (i, a)], inp["c%d_kc%d" % (i, a)] = cos[a], s2[a], kc[a]
        inp["c%d_Y" % i], inp["c%d_Z" % i], inp["c%d_co" % i], inp["c%d_xr" % i] = Y, Z, co, np.array(xr)
        if wk == "plane":
            inp["c%d_w" % i] = rng.random((1, N, N))
        elif wk == "dense":
            inp["c%d_w" % i] = rng.random((N, N, N))
    np.savez(tmp_path / "in.npz", **inp)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for tag, lib in (("fast", None), ("ref", _build.LIB_REFORDER)):
        env = dict(os.environ)
        env.pop("SB200_LIB", None)
        if lib:
            env["SB200_LIB"] = lib
        outp = str(tmp_path / (tag + ".npz"))
        script = _VARIANT_SCRIPT % dict(root=root, inp=str(tmp_path / "in.npz"), outp=outp)
        subprocess.run([sys.executable, "-c", script], env=env, check=True, timeout=600)
        res[tag] = np.load(outp)
    n_finite = 0
    for i in range(len(cases)):
        f, r = res["fast"], res["ref"]
        assert np.array_equal(f["lo%d" % i].view(np.int64), r["lo%d" % i].view(np.int64)), cases[i]
        assert np.array_equal(f["hi%d" % i].view(np.int64), r["hi%d" % i].view(np.int64)), cases[i]
        assert np.array_equal(f["nan%d" % i] > 0, r["nan%d" % i] > 0), cases[i]
        ok = r["nan%d" % i] == 0
        assert np.array_equal(np.isnan(f["S%d" % i]), np.isnan(r["S%d" % i]))
        assert np.max(relerr(f["S%d" % i][ok], r["S%d" % i][ok])) <= 1e-13, cases[i]
        n_finite += int(ok.sum())
    assert n_finite >= 300


def test_group_velocity_is_numpy_gradient(sb):
    """sb200_group_velocity against numpy on the SAME omega: gradient components, |vg|, vph and k bitwise
    (dispersion.py:208-220: np.gradient(omega, *dks, edge_order=2), sqrt of the sum of squares, omega/k)."""
    for dim, N, Y, Z, fam in ((3, 23, 1.0, 1.0, "StencilFree3D"), (3, 9, 1.4, 0.6, "StencilFree3D"),
                              (2, 57, 1.0, 1.0, "StencilFree2D"), (2, 3, 2.5, 1.0, "StencilFree2D")):
        grid = O.Grid(dim, N, Y, Z)
        p = np.zeros(len(O.parameter_names(fam)))
        p[0], p[-1] = 0.35, -0.04
        co = O.coefficients(fam, p)
        dks = [kc.ravel()[1] for kc in grid.kcomp]
        with ctx_from_grid(sb, grid) as ctx:
            f = ctx.group_velocity(co, dks)
            om = ctx.export(co, 0)[0]
        np.testing.assert_array_equal(f["omega"], om)
        vgs = np.gradient(om, *dks, edge_order=2)
        for a in range(dim):
            np.testing.assert_array_equal(f["vgc"][a], vgs[a])
        np.testing.assert_array_equal(f["vg"], np.sqrt(sum(v ** 2 for v in vgs)))
        np.testing.assert_array_equal(f["k"], np.broadcast_to(grid.k, grid.shape))
        with np.errstate(invalid="ignore", divide="ignore"):
            np.testing.assert_array_equal(f["vph"], om / grid.k)
        assert f["nan"] == 0
    with ctx_from_grid(sb, O.Grid(2, 2)) as ctx, pytest.raises(sb.StencilB200Error):
        ctx.group_velocity(O.coefficients("StencilFree2D", [0.3, 0, 0, 0, 0]), [1.0, 1.0])
