"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    python oracle/make_golden.py

It imports the reference through oracle/ref_shim.py (NumPy-alias shim only),
evaluates a fixed list of cases with the reference's own Dispersion2D/3D objects and
Optimize driver, and writes the inputs (including the reference's own per-axis tables) and
outputs.  The fixtures pin the oracle (oracle/np_oracle.py, oracle/oracle_dispersion.c) and
are what the GPU parity tests compare against on the GPU box, where the reference is absent.
"""
import argparse
import contextlib
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
NSAMPLES = 96


def sample_indices(case_id, size):
    rng = np.random.default_rng(1000 + case_id)
    idx = rng.integers(0, size, NSAMPLES)
    idx[0], idx[1] = 0, size - 1  # always the origin (omega = 0) and the far corner
    return idx


def case_list():
    """(stencil, dim, N, Y, Z, weight, weight_params, dt_multiplier, params, tag)"""
    C = []
    add = lambda *a: C.append(a)
    # the reference's own smoke-test points (tests/test_dispersion.py:25-61)
    add("StencilDivFree2D", 2, 100, 1, 1, "equal", (), 1.0, [0.3, -0.10, -0.60], "reftest-2d")
    add("StencilFree2D", 2, 100, 1, 1, "equal", (), 1.0, [0.3, 0.3, -0.3, 0.3, 0.7], "reftest-2d-bad")
    add("StencilDivFree3D", 3, 100, 1, 1, "equal", (), 1.0, [0.3, -0.10, -0.60, 0.2], "reftest-3d")
    # config 2: calculate_omega.py --lehe --params 0.96 (2D N=200 literal README command; 3D form)
    dx_lehe = 0.25 * (1 - 1 / 0.96 ** 2 * np.sin(np.pi * 0.96 / 2.0))
    add("StencilSymmetric2D", 2, 200, 1, 1, "equal", (), 1.0, [0.96, 0.125, dx_lehe, 0.0], "lehe-2d-200")
    add("StencilSymmetric3D", 3, 64, 1, 1, "equal", (), 1.0, [0.96, .125, .125, .125, dx_lehe, 0, 0], "lehe-3d-64")
    add("StencilSymmetric3D", 3, 256, 1, 1, "equal", (), 1.0, [0.96, .125, .125, .125, dx_lehe, 0, 0], "lehe-3d-256")
    # on the CFL limit: the --div-free optimum (max s = 1 - 2e-11) and the default-run optimum
    add("StencilIsotropicDivFree2D", 2, 200, 1, 1, "equal", (), 1.0,
        [0.6718000000000016, -0.013484133612260314], "divfree-opt-2d")
    add("StencilSymmetric2D", 2, 200, 1, 1, "equal", (), 1.0,
        [0.6857978707687641, 0.10963493621568149, -0.12476172846832859, -0.12479339137808687], "default-opt-2d")
    add("StencilSymmetric2D", 2, 50, 1, 1, "equal", (), 1.0,
        [0.6857978707687641, 0.10963493621568149, -0.12476172846832859, -0.12479339137808687], "default-opt-2d-low")
    # Yee at exactly its CFL time step (max s == 1.0): dt filled in below from the reference's dt_ok
    add("StencilSymmetric2D", 2, 100, 1, 1, "equal", (), 1.0, ["CFL", 0.0, 0.0, 0.0], "yee-2d-cfl")
    add("StencilSymmetric3D", 3, 48, 1, 1, "equal", (), 1.0, ["CFL", 0, 0, 0, 0, 0, 0], "yee-3d-cfl")
    add("StencilSymmetric3D", 3, 40, 1.5, 2.5, "equal", (), 1.0, ["CFL", 0, 0, 0, 0, 0, 0], "yee-3d-cfl-aniso")
    # config 4's centre point and neighbours (StencilFree3D, 10 free parameters)
    xc = np.array([0.5, .02, .03, .01, .02, .03, .01, -.02, -.03, -.01])
    add("StencilFree3D", 3, 64, 1, 1, "equal", (), 1.0, xc, "config4-centre")
    rng = np.random.default_rng(0)
    for i in range(6):
        add("StencilFree3D", 3, 40, 1, 1, "equal", (), 1.0, xc + 1e-3 * rng.standard_normal(10), "config4-rand%d" % i)
    # config 3's family: StencilIsotropicDivFree3D (dt, deltax), some gated
    for dt, dlx in [(0.6718, -0.02), (0.70, -0.0135), (0.9, -0.05), (0.99, 0.2), (0.75, -0.9), (0.68, 0.1)]:
        add("StencilIsotropicDivFree3D", 3, 48, 1, 1, "equal", (), 1.0, [dt, dlx], "config3-scan")
    # every stencil family, anisotropic cells, every weight kind, dt_multiplier != 1
    rng = np.random.default_rng(7)
    fams2 = ["StencilFree2D", "StencilDivFree2D", "StencilSymmetric2D", "StencilIsotropic2D",
             "StencilIsotropicDivFree2D"]
    fams3 = ["StencilFree3D", "StencilDivFree3D", "StencilSymmetric3D", "StencilCylinder3D",
             "StencilCylinderDivFree3D", "StencilIsotropic3D", "StencilIsotropicDivFree3D"]
    weights = [("equal", ()), ("xaxis", ()), ("xaxis_soft", ()), ("xaxis_soft", (0.35,))]
    es = ref_shim.load_reference()
    k = 0
    for fam in fams2 + fams3:
        dim = 2 if fam.endswith("2D") else 3
        dof = getattr(es, fam)().dof
        for rep in range(4):
            N = [50, 37, 64, 33][rep] if dim == 2 else [24, 17, 32, 20][rep]
            Y = [1, 1.25, 0.8, 2.0][rep]
            Z = [1, 1, 1.6, 0.75][rep]
            wk, wp = weights[(k + rep) % 4]
            p = 0.08 * rng.standard_normal(dof)
            p[0] = [0.3, 0.45, 0.2, 0.55][rep]
            mult = [1.0, 1.0, 0.95, 1.0][rep]
            add(fam, dim, N, Y, Z, wk, wp, mult, p, "family")
        k += 1
    # infeasible / degenerate inputs
    add("StencilFree2D", 2, 50, 1, 1, "equal", (), 1.0, [1.5, 0.0, 0.0, 0.0, 0.0], "dt-too-large")
    add("StencilFree3D", 3, 24, 1, 1, "equal", (), 1.0, [2.0, 0, 0, 0, 0, 0, 0, 0, 0, 0], "dt-too-large-3d")
    add("StencilFree2D", 2, 50, 1, 1, "equal", (), 1.0, [-0.1, 0.0, 0.0, 0.0, 0.0], "dt-negative")
    add("StencilFree2D", 2, 50, 1, 1, "equal", (), 1.0, [0.0, 0.0, 0.0, 0.0, 0.0], "dt-zero")
    add("StencilFree2D", 2, 50, 1, 1, "equal", (), 1.0, [0.5, np.nan, 0.0, 0.0, 0.0], "nan-param")
    add("StencilFree3D", 3, 24, 1, 1, "xaxis", (), 1.0, [0.4, .6, -.5, .4, .3, -.7, .2, .24, -.9, .1], "bad-3d")
    add("StencilSymmetric3D", 3, 24, 1, 1, "equal", (), 1.0, [0.4, 0, 0, 0, 0.3, 0.3, 0.3], "corner-negative")
    return C


def run_cases():
    es = ref_shim.load_reference()
    cases = case_list()
    rec = {k: [] for k in ("stencil", "dim", "N", "Y", "Z", "wkind", "wparam", "dtmul", "nparams", "params",
                           "coeffs", "tag", "stencil_ok", "dt_ok", "norm", "gated", "Amin", "Amax", "S",
                           "omega_samples", "grid_id")}
    grids, grid_keys, full_omega = [], {}, {}
    for cid, (fam, dim, N, Y, Z, wk, wp, mult, p, tag) in enumerate(cases):
        st = getattr(es, fam)()
        d = es.Dispersion2D(Y, N, st) if dim == 2 else es.Dispersion3D(Y, Z, N, st)
        d.dt_multiplier = mult
        if wk != "equal":
            es.weight_functions[wk](*wp)(d)
        p = list(p)
        if isinstance(p[0], str):  # Yee-style: dt := the CFL limit the reference itself reports
            p[0] = 0
            p[0] = np.asarray(d.dt_ok(p)).item()
        p = np.asarray(p, dtype=float)
        key = (dim, N, float(Y), float(Z))
        if key not in grid_keys:
            grid_keys[key] = len(grids)
            if dim == 2:
                cos = [d.coskappax, d.coskappay]; s2 = [d.sx2, d.sy2]; kc = [d.kx, d.ky]
            else:
                cos = [d.coskappax, d.coskappay, d.coskappaz]; s2 = [d.sx2, d.sy2, d.sz2]; kc = [d.kx, d.ky, d.kz]
            grids.append(dict(key=key, cos=np.stack([a.ravel() for a in cos]), s2=np.stack([a.ravel() for a in s2]),
                              kcomp=np.stack([a.ravel() for a in kc])))
        with np.errstate(invalid="ignore", divide="ignore"):
            sok = np.asarray(d.stencil_ok(p), dtype=float).ravel()[0]
            dtok = np.asarray(d.dt_ok(p), dtype=float).ravel()[0]
            nrm = float(d.norm(p))
            om = d.omega(p) if not np.any(np.isnan(p)) else None
            if np.any(np.isnan(p)):
                Amin = Amax = np.nan
                coeffs = np.full(len(st.Coefficients._fields), np.nan)
            else:
                A = np.asarray(d.sqrtarg)
                Amin, Amax = np.min(A), np.max(A)
                coeffs = np.array([d.coefficients[f][0] for f in st.Coefficients._fields])
        if om is None:
            S, oms, gated = np.nan, np.full(NSAMPLES, np.nan), 1
        else:
            S = float((d.w * (om - d.k) ** 2).sum())
            oms, gated = om.ravel()[sample_indices(cid, om.size)], 0
            if om.size <= 64 ** 3 and tag in ("lehe-3d-64", "reftest-2d", "divfree-opt-2d", "yee-2d-cfl", "yee-3d-cfl"):
                full_omega["omega_full_%d" % cid] = om
        pad = lambda v, n: np.concatenate([v, np.full(n - len(v), np.nan)])
        for k, v in dict(stencil=fam, dim=dim, N=N, Y=Y, Z=Z, wkind=wk, wparam=(wp[0] if wp else np.nan), dtmul=mult,
                         nparams=len(p), params=pad(p, 10), coeffs=pad(coeffs, 13), tag=tag, stencil_ok=sok,
                         dt_ok=dtok, norm=nrm, gated=gated, Amin=Amin, Amax=Amax, S=S, omega_samples=oms,
                         grid_id=grid_keys[key]).items():
            rec[k].append(v)
        print("%3d %-28s dim=%d N=%3d Y=%.2f Z=%.2f w=%-10s ok=% .6g dt_ok=% .6g norm=%.17g"
              % (cid, fam, dim, N, Y, Z, wk, sok, dtok, nrm))
    out = {k: np.array(v) for k, v in rec.items()}
    for gi, g in enumerate(grids):
        out["grid%d_key" % gi] = np.array(g["key"])
        out["grid%d_cos" % gi] = g["cos"]
        out["grid%d_s2" % gi] = g["s2"]
        out["grid%d_kcomp" % gi] = g["kcomp"]
    out["ngrids"] = np.array(len(grids))
    out.update(full_omega)
    return out


def run_optimisations():
    """End-to-end anchors: the reference's Optimize driver on BASELINE.json configs 1 and 3 (2D form)."""
    es = ref_shim.load_reference()
    base = dict(N=3, scan_dt=True, Ngrid_low=50, Ngrid_high=200, dim=2, div_free=False, symmetric_axes=0,
                symmetric_beta=True, Y=1.0, Z=1.0, dt_multiplier=1.0, deltaxrange=[-1, 0.25],
                deltayrange=[-1, 0.25], deltazrange=[-1, 0.25], betaxyrange=[-1, 1], betaxzrange=[-1, 1],
                betayxrange=[-1, 1], betayzrange=[-1, 1], betazxrange=[-1, 1], betazyrange=[-1, 1],
                dtrange=[0.1, 1], weight="equal", weight_params=[], singlecore=True)
    runs = {"default": {}, "divfree": dict(div_free=True, dtrange=[0.6718, 1.0]),
            "xaxis_soft": dict(weight="xaxis_soft", weight_params=[0.3], N=2),
            "divfree3d_small": dict(dim=3, div_free=True, dtrange=[0.55, 1.0], Ngrid_low=16, Ngrid_high=24, N=2)}
    out = {}
    for name, over in runs.items():
        args = types.SimpleNamespace(**{**base, **over})
        log = io.StringIO()
        with contextlib.redirect_stdout(log):
            opt = es.Optimize(args)
            x, fmin = opt.optimize()
        lines = [l for l in log.getvalue().splitlines() if l.startswith("x0=")]
        out[name + "_x"] = np.asarray(x, dtype=float)
        out[name + "_norm"] = np.asarray(float(fmin))
        out[name + "_nstarts"] = np.asarray(len(lines))
        out[name + "_stencil"] = np.asarray(opt.stencil.__class__.__name__)
        out[name + "_stencil_ok"] = np.asarray(opt.dispersion_high.stencil_ok(x), dtype=float).ravel()
        out[name + "_dt_ok"] = np.asarray(opt.dispersion_high.dt_ok(x), dtype=float).ravel()
        print(name, opt.stencil.__class__.__name__, x, fmin)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-optimise", action="store_true")
    a = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "objective_cases.npz"), **run_cases())
    if not a.skip_optimise:
        np.savez_compressed(os.path.join(OUT, "optimise_runs.npz"), **run_optimisations())


if __name__ == "__main__":
    main()
