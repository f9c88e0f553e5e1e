"""GPU tests of the ABI-v2 entry points (through the C-ABI): segmented launches, strided k-slab plane lists, the
start-grid screening kernel variant, multi-GPU device groups, the device-side slab combine, the chunked export
copy, and the lock-stepped multi-start scan built on top of them."""
import numpy as np
import pytest

import np_oracle as O
from conftest import relerr

pytestmark = pytest.mark.gpu

XC = np.array([0.5, .02, .03, .01, .02, .03, .01, -.02, -.03, -.01])


@pytest.fixture(scope="module")
def sb():
    import optimize_stencil_b200 as m
    assert m.device_count() >= 1, "no CUDA device visible"
    return m


def ctx_from_grid(sb, grid, **kw):
    cos, s2, kc = grid.flat_tables()
    return sb.Context(grid.dim, cos, s2, kc, Y=grid.Y, Z=grid.Z, **kw)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.int64)


def same_result(a, b):
    """(S, Amin, Amax, nan) tuples equal BIT FOR BIT (NaNs: same positions)."""
    for x, y in zip(a[:3], b[:3]):
        x, y = np.asarray(x), np.asarray(y)
        if not np.array_equal(np.isnan(x), np.isnan(y)):
            return False
        ok = ~np.isnan(x)
        if not np.array_equal(bits(x[ok]), bits(y[ok])):
            return False
    return np.array_equal(np.asarray(a[3]) > 0, np.asarray(b[3]) > 0)


def mixed_candidates(rng, n, fam="StencilFree3D"):
    """Feasible rows near x_c, rows beyond the CFL limit (NaN sums), a negative-sqrtarg row."""
    dof = len(O.parameter_names(fam))
    P = np.tile(XC[:dof], (n, 1)) + 1e-3 * rng.standard_normal((n, dof))
    P[::5, 0] = rng.uniform(1.2, 2.5, size=len(P[::5]))
    if fam == "StencilFree3D" and n > 3:
        P[3] = [0.4, .6, -.5, .4, .3, -.7, .2, .24, -.9, .1]
    return np.stack([O.coefficients(fam, p) for p in P])


@pytest.mark.parametrize("dim,N,Y,Z,fam,seg", [(3, 40, 1.0, 1.0, "StencilFree3D", 11), (3, 24, 1.3, 0.8, "StencilFree3D", 3),
                                                (3, 33, 1.0, 1.0, "StencilFree3D", 1), (3, 33, 1.0, 1.0, "StencilFree3D", 2),
                                                (2, 50, 1.0, 1.0, "StencilFree2D", 5), (2, 200, 1.25, 1.0, "StencilFree2D", 5),
                                                (3, 128, 1.0, 1.0, "StencilFree3D", 3)])
def test_segments_are_bit_identical_to_their_own_launches(sb, dim, N, Y, Z, fam, seg):
    """The property the lock-stepped SLSQP scan rests on: every segment of a segmented launch gives the bits of a
    stand-alone launch of that segment, whatever else shares the launch (include/stencil_b200.h, seg_size)."""
    rng = np.random.default_rng(dim * 1000 + N + seg)
    nseg = 9
    co = mixed_candidates(rng, nseg * seg, fam)
    with ctx_from_grid(sb, O.Grid(dim, N, Y, Z)) as ctx:
        whole = ctx.objective_batch(co, seg_size=seg)
        for s in range(nseg):
            sl = slice(s * seg, (s + 1) * seg)
            alone = ctx.objective_batch(co[sl])
            assert same_result([w[sl] for w in whole], alone), (s, [w[sl] for w in whole], alone)
        # ... and in a different company: the segments reversed and only five of them
        order = [4, 2, 8, 0, 6]
        part = ctx.objective_batch(np.concatenate([co[s * seg:(s + 1) * seg] for s in order]), seg_size=seg)
        for n, s in enumerate(order):
            sl = slice(s * seg, (s + 1) * seg)
            assert same_result([p[n * seg:(n + 1) * seg] for p in part], [w[sl] for w in whole]), s


def test_segment_argument_checks(sb):
    with ctx_from_grid(sb, O.Grid(3, 12)) as ctx:
        co = mixed_candidates(np.random.default_rng(0), 6)
        with pytest.raises(sb.StencilB200Error, match="multiple of seg_size"):
            ctx.objective_batch(co, seg_size=4)
        with pytest.raises(sb.StencilB200Error, match="plane list"):
            ctx.objective_batch(co, x_begin=3, x_end=40)
    with ctx_from_grid(sb, O.Grid(2, 12)) as ctx:
        co = mixed_candidates(np.random.default_rng(0), 2, "StencilFree2D")
        with pytest.raises(sb.StencilB200Error, match="dim 3 only"):
            ctx.objective_batch(co, x_begin=0, x_end=12, x_stride=2)


def test_strided_plane_lists_partition_the_grid(sb):
    """Round-robin k-slabs: planes g, g+G, ... against the oracle on exactly those planes, and their combination
    against the whole grid (sum to 1e-13, extrema bit for bit)."""
    N, G = 45, 4
    grid = O.Grid(3, N, 1.2, 0.9)
    co = mixed_candidates(np.random.default_rng(5), 7)
    with ctx_from_grid(sb, grid) as ctx:
        full = ctx.objective_batch(co)
        parts = [ctx.objective_batch(co, x_begin=g, x_end=N, x_stride=G) for g in range(G)]
        for g in range(G):
            sub = O.Grid(3, N, 1.2, 0.9, xslab=(g, N, G))
            for b in range(len(co)):
                S, amin, amax, nn = O.device_contract(sub, co[b])
                assert bits(parts[g][1][b]) == bits(amin if amin < 0 else 0.0) or np.isnan(amin)
                assert bits(parts[g][2][b]) == bits(max(amax, 0.0)) or np.isnan(amax)
                if nn == 0:
                    assert relerr(parts[g][0][b], S) <= 1e-12
                assert (parts[g][3][b] > 0) == (nn > 0)
        ok = ~np.isnan(full[0])
        assert np.max(relerr(sum(p[0] for p in parts)[ok], full[0][ok])) <= 1e-13
        np.testing.assert_array_equal(np.maximum.reduce([p[2] for p in parts]), full[2])
        np.testing.assert_array_equal(np.minimum.reduce([p[1] for p in parts]), full[1])


def test_screening_variant_gives_the_same_results(sb):
    """SB200_F_SCAN skips warps whose k-points all have s > 1; results must not change -- also for a start-grid-like
    batch (config 3: mostly beyond the CFL limit) and for the automatic choice."""
    st_n = 24
    dts = np.linspace(0.6718, 1.0, st_n + 2)[1:-1]
    dxs = np.linspace(-1, 0.25, st_n + 2)[1:-1]
    co = np.stack([O.coefficients("StencilIsotropicDivFree3D", (t, x)) for t in dts for x in dxs])
    with ctx_from_grid(sb, O.Grid(3, 64)) as ctx:
        plain = ctx.objective_batch(co, fast=True)                   # the screening variant uses the fast series
        assert np.mean(np.isnan(plain[0])) > 0.5                     # a screening-like batch indeed
        assert same_result(plain, ctx.objective_batch(co, scan=True))
        assert same_result(plain, ctx.objective_batch(co, scan="auto"))
        few = mixed_candidates(np.random.default_rng(2), 97)
        assert same_result(ctx.objective_batch(few, fast=True), ctx.objective_batch(few, scan=True))
    with ctx_from_grid(sb, O.Grid(2, 90, 1.5)) as ctx:
        co2 = mixed_candidates(np.random.default_rng(3), 130, "StencilFree2D")
        assert same_result(ctx.objective_batch(co2, fast=True), ctx.objective_batch(co2, scan=True))


@pytest.mark.parametrize("dim,N,Y,Z", [(3, 40, 1.0, 1.0), (3, 33, 1.3, 0.8), (2, 64, 1.0, 1.0), (2, 50, 1.25, 1.0)])
def test_chunks_of_hopeless_candidates_take_the_extrema_only_sweep(sb, dim, N, Y, Z):
    """Chunks whose candidates all have s > 1 or sqrtarg < 0 at a corner of the grid skip omega (the sum is NaN
    anyway) and only look for min / max of sqrtarg: NaN sums, extrema bit for bit -- with the maximum filter, the
    minimum filter (large negative regions) and its fallback (barely negative values) all in play."""
    rng = np.random.default_rng(dim * 100 + N)
    fam = "StencilFree3D" if dim == 3 else "StencilFree2D"
    dof = len(O.parameter_names(fam))
    rows = []
    for _ in range(24):                                  # beyond the CFL limit: dt 1.1 .. 3
        p = np.zeros(dof); p[0] = rng.uniform(1.1, 3.0); p[1:] = 0.02 * rng.standard_normal(dof - 1); rows.append(p)
    for _ in range(24):                                  # large negative regions: wild betas / deltas
        p = 0.6 * rng.standard_normal(dof); p[0] = rng.uniform(0.2, 0.9); rows.append(p)
    for eps in (1e-3, 1e-9, 1e-14):                      # barely negative at the far corner: delta = (1 + eps) / 4 (Yee-like)
        p = np.zeros(dof); p[0] = 0.3; p[-dim:] = 0.25 * (1 + eps); rows.append(p)
    co = [O.coefficients(fam, p) for p in rows]
    if dim == 3:
        # what an SLSQP line search throws at the delta = 1/4 bound of the div-free isotropic family (config 3): corners
        # 0 / +-1e-12 / -2 / -6, four fifths of the grid negative, the maximum 1/4 on a surface inside the grid
        for dt, d in ((0.6718, 0.25), (0.6718000001, 0.25), (0.7, 0.25 - 1e-9), (0.6718, 0.26), (0.9, 0.3), (0.5, 0.2501)):
            co.append(O.coefficients("StencilIsotropicDivFree3D", [dt, d]))
    while True:                                          # sqrtarg = 0 everywhere (dt = 0: not an s > 1 case); also pads
        co.append(np.zeros(len(co[0])))                  # the batch to whole segments of 3
        if len(co) % 3 == 0:
            break
    co = np.stack(co)
    grid = O.Grid(dim, N, Y, Z)
    with ctx_from_grid(sb, grid) as ctx:
        for kw in (dict(), dict(seg_size=3), dict(fast=True)):
            S, lo, hi, nan = ctx.objective_batch(co, **kw)
            dead = 0
            for b in range(len(co)):
                Sr, amin, amax, nn = O.device_contract(grid, co[b])
                assert bits(lo[b]) == bits(amin if amin < 0 else 0.0), (b, kw, lo[b], amin)
                assert bits(hi[b]) == bits(max(amax, 0.0)), (b, kw, hi[b], amax)
                assert (nan[b] > 0) == (nn > 0) and np.isnan(S[b]) == (nn > 0), (b, kw)
                if nn == 0:
                    assert relerr(S[b], Sr) <= 1e-12
                dead += nn > 0
            assert dead >= 40
        # a hopeless candidate next to healthy ones in the same chunk: the full sweep serves both
        mix = np.stack([co[0], O.coefficients(fam, XC[:dof]), co[30], O.coefficients(fam, XC[:dof] * 0.9)])
        S, lo, hi, nan = ctx.objective_batch(mix)
        for b in range(4):
            Sr, amin, amax, nn = O.device_contract(grid, mix[b])
            assert bits(hi[b]) == bits(max(amax, 0.0)) and bits(lo[b]) == bits(amin if amin < 0 else 0.0)
            assert np.isnan(S[b]) == (nn > 0) and (nn > 0 or relerr(S[b], Sr) <= 1e-12)
        # omega = (2/dt) asin(dt sqrt(A)) is even in dt, bit for bit (dispersion.py:191): so is the device's
        flipped = mix.copy()
        flipped[:, 0] = -flipped[:, 0]
        assert same_result(ctx.objective_batch(flipped), (S, lo, hi, nan))


def test_fast_and_precise_series(sb):
    """SB200_F_FAST trades 2.3e-16 for 3.7e-15 in asin: extrema and NaN flags identical, sums within 1e-14 of each
    other and both inside the 1e-12 bar against the oracle; the automatic choice goes by segment size."""
    grid = O.Grid(3, 40)
    co = mixed_candidates(np.random.default_rng(17), 21)
    with ctx_from_grid(sb, grid) as ctx:
        precise, fast = ctx.objective_batch(co, fast=False), ctx.objective_batch(co, fast=True)
        assert same_result(precise, ctx.objective_batch(co))          # small launch: precise by default
        np.testing.assert_array_equal(precise[1], fast[1]); np.testing.assert_array_equal(precise[2], fast[2])
        np.testing.assert_array_equal(precise[3] > 0, fast[3] > 0)
        ok = ~np.isnan(precise[0])
        assert np.max(relerr(fast[0][ok], precise[0][ok])) <= 1e-14
        worst = {"precise": 0.0, "fast": 0.0}
        for b in np.flatnonzero(ok):
            S = O.device_contract(grid, co[b])[0]
            worst["precise"] = max(worst["precise"], relerr(precise[0][b], S))
            worst["fast"] = max(worst["fast"], relerr(fast[0][b], S))
        assert worst["precise"] <= 2e-15 and worst["fast"] <= 1e-14, worst


def test_combine_kernel_matches_the_host_combine(sb):
    import torch
    from optimize_stencil_b200 import sharding
    rng = np.random.default_rng(9)
    world, B = 5, 37
    parts = rng.standard_normal((world, 4, B))
    parts[:, 3] = rng.integers(0, 3, (world, B))
    parts[2, 1, 4] = np.nan
    parts[1, 2, 9] = np.nan
    parts[3, 0, 11] = np.nan
    want = sharding.combine_slab_partials(torch.from_numpy(parts)).numpy()
    with ctx_from_grid(sb, O.Grid(3, 8)) as ctx:
        d = torch.from_numpy(parts).cuda()
        got = sharding.device_combiner(ctx)(d).cpu().numpy()
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
    np.testing.assert_array_equal(bits(got[~np.isnan(got)]), bits(want[~np.isnan(want)]))


def test_device_and_host_entry_points_share_scratch_safely(sb):
    """a host-path call right after an asynchronous *_device call on another stream must wait for it."""
    import torch
    grid = O.Grid(3, 48)
    co = mixed_candidates(np.random.default_rng(4), 64)
    with ctx_from_grid(sb, grid) as ctx:
        want = ctx.objective_batch(co)
        dco = torch.from_numpy(co).cuda()
        out = torch.empty((4, len(co)), dtype=torch.float64, device="cuda")
        side = torch.cuda.Stream()
        for _ in range(5):
            with torch.cuda.stream(side):
                ctx.objective_batch_device(dco.data_ptr(), len(co), out.data_ptr(), stream=side.cuda_stream)
            got_host = ctx.objective_batch(co[::-1].copy())             # no host synchronisation in between
            side.synchronize()
            got_dev = out.cpu().numpy()
            assert same_result(want, (got_dev[0], got_dev[1], got_dev[2], got_dev[3]))
            assert same_result([w[::-1] for w in want], got_host)


def test_export_chunked_copy_is_exact_and_reports_its_rate(sb):
    """A field larger than the staging chunk (16 MiB) through the double-buffered pinned copy."""
    N = 160                                                            # 160^3 doubles = 31 MiB: two and a bit chunks
    grid = O.Grid(3, N)
    co = O.coefficients("StencilFree3D", XC)
    with ctx_from_grid(sb, grid) as ctx:
        A, amin, amax, nn = ctx.export(co, 1)
        np.testing.assert_array_equal(A, O.sqrtarg(grid, co))
        nbytes, secs = ctx.last_export_stats
        assert nbytes == A.nbytes and secs > 0
        f = ctx.group_velocity(co, (grid.kcomp[0].ravel()[1], grid.kcomp[1].ravel()[1], grid.kcomp[2].ravel()[1]))
        om = ctx.export(co, 0)[0]
        np.testing.assert_array_equal(f["omega"], om)
        ref = np.gradient(om, *[k.ravel()[1] for k in grid.kcomp], edge_order=2)
        for a in range(3):
            np.testing.assert_array_equal(f["vgc"][a], ref[a])


def need_devices(sb, n):
    if sb.device_count() < n:
        pytest.skip("needs %d CUDA devices, %d visible" % (n, sb.device_count()))


def test_device_group_matches_one_device(sb):
    """sb200_group_*: the batch sharded over the GPUs of this process (SURVEY.md 8b.1 'device list', 8e)."""
    need_devices(sb, 2)
    G = min(sb.device_count(), 8)
    grid = O.Grid(3, 72, 1.1, 0.95)
    rng = np.random.default_rng(21)
    co = mixed_candidates(rng, 203)
    with ctx_from_grid(sb, grid) as one, ctx_from_grid(sb, grid, devices=list(range(G))) as grp:
        want = one.objective_batch(co)
        # candidates: extrema and NaN flags exact, sums to the last bits (the tiling follows the shard size)
        got = grp.objective_batch(co, shard="candidates")
        assert grp.last_shard == ("candidates", G)
        np.testing.assert_array_equal(got[1], want[1]); np.testing.assert_array_equal(got[2], want[2])
        np.testing.assert_array_equal(got[3] > 0, want[3] > 0)
        ok = ~np.isnan(want[0])
        assert np.max(relerr(got[0][ok], want[0][ok])) <= 1e-13
        # segmented: bit-identical to one device, whatever the number of devices
        seg = 7
        assert same_result(grp.objective_batch(co, seg_size=seg, shard="candidates"), one.objective_batch(co, seg_size=seg))
        # slabs, both dealings
        for mode in ("slabs", "slabs-contiguous"):
            got = grp.objective_batch(co[:9], shard=mode)
            assert grp.last_shard == (mode, G)
            np.testing.assert_array_equal(got[1], want[1][:9]); np.testing.assert_array_equal(got[2], want[2][:9])
            ok = ~np.isnan(want[0][:9])
            assert np.array_equal(np.isnan(got[0]), ~ok) and np.max(relerr(got[0][ok], want[0][:9][ok])) <= 1e-13
        # weights travel to every device
        w = O.weight(grid, "xaxis_soft", (0.3,))
        one.set_weight(w); grp.set_weight(w)
        a, b = one.objective_batch(co[:40]), grp.objective_batch(co[:40], shard="candidates")
        ok = ~np.isnan(a[0])
        assert np.max(relerr(b[0][ok], a[0][ok])) <= 1e-13
        # small batches stay on one device in auto mode; exports use the first device
        grp.objective_batch(co[:3])
        assert grp.last_shard == ("none", 1)
        np.testing.assert_array_equal(grp.export(co[0], 1)[0], one.export(co[0], 1)[0])
        assert grp.launch_count > one.launch_count * 0


def test_dispersion_uses_all_devices_for_large_batches(sb):
    need_devices(sb, 2)
    from optimize_stencil_b200 import extended_stencil as es
    d1 = es.Dispersion3D(1.0, 1.0, 96, es.StencilFree3D())
    d1.devices = [0]
    dn = es.Dispersion3D(1.0, 1.0, 96, es.StencilFree3D())
    dn.auto_multi_evals = 1e6                                           # bring the other GPUs up for this test's size
    P = XC + 1e-3 * np.random.default_rng(0).standard_normal((512, 10))
    a, b = d1.evaluate_batch(P), dn.evaluate_batch(P)
    assert len(dn._context().devices) == sb.device_count() and dn._context().last_shard[0] == "candidates"
    np.testing.assert_array_equal(a["stencil_ok"], b["stencil_ok"])
    np.testing.assert_array_equal(a["dt_ok"], b["dt_ok"])
    assert np.max(relerr(b["norm"], a["norm"])) <= 1e-13


@pytest.mark.parametrize("over", [dict(div_free=True, dtrange=[0.6718, 1.0]), {},
                                  dict(dim=3, div_free=True, dtrange=[0.55, 1.0], Ngrid_low=32, Ngrid_high=48, N=4)])
def test_lockstep_scan_is_bit_identical_to_the_sequential_scan(sb, over, capsys):
    """Optimize.optimize(): all searches concurrently, one segmented launch per round, against the searches one
    after the other -- every printed line (x0, x, norm of every start) must be identical."""
    from optimize_stencil_b200 import extended_stencil as es
    from test_dropin_api import default_args
    runs = {}
    for flag in (False, True):
        es.Optimize.lockstep = flag
        try:
            opt = es.Optimize(default_args(**over))
            x, f = opt.optimize()
        finally:
            es.Optimize.lockstep = True
        runs[flag] = (x, f, capsys.readouterr().out, opt.dispersion.stats["evaluated"], opt.scan_stats)
    assert runs[True][2] == runs[False][2]
    assert np.array_equal(runs[True][0], runs[False][0]) and runs[True][1] == runs[False][1]
    assert runs[True][3] == runs[False][3]                              # the same points were evaluated
    assert runs[True][4]["launches"] < runs[True][3] / 3                # ... in far fewer launches


def test_worker_processes_served_by_the_device_owner(sb, capsys):
    """SciPy-only worker processes (spawned; no CUDA in them) served by this process, which owns all visible GPUs:
    the in-process result bit for bit."""
    from optimize_stencil_b200 import extended_stencil as es
    from test_dropin_api import default_args
    over = dict(dim=3, div_free=True, dtrange=[0.55, 1.0], Ngrid_low=24, Ngrid_high=24, N=5)
    x1, f1 = es.Optimize(default_args(**over)).optimize()
    out1 = capsys.readouterr().out
    opt = es.Optimize(default_args(workers=3, singlecore=False, **over))
    x2, f2 = opt.optimize()
    out2 = capsys.readouterr().out
    assert out1 == out2 and f1 == f2 and np.array_equal(x1, x2)
    assert opt.scan_stats["workers"] == 3 and opt.scan_stats["devices"] == sb.device_count()
    assert opt.scan_stats["launches"] < opt.scan_stats["rows"] / 3


def test_fp64_peak_reaches_the_pipe_rate(sb):
    """The measured roofline denominator: within 4 % of one warp instruction per 2 cycles per sub-partition
    (148 SMs x 64 lanes x 2 flop x SM clock) at whatever clock the GPU runs the microbenchmark."""
    peak = sb.fp64_peak_tflops(0, 0.5)
    assert 30.0 < peak < 39.0, peak
