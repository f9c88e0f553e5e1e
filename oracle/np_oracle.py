"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference objective.

This is the CPU oracle for the hot path named in BASELINE.json (`north_star`):
the objective that `Optimize.optimize()` minimises.  It restates, pass by
pass and in the reference's operation order, what
/root/reference/extended_stencil/{dispersion,stencil,weight}.py compute, so
that (a) it can be checked bit-for-bit against the unmodified reference where
that is importable (tests/test_oracle_vs_reference.py, oracle/make_golden.py)
and (b) it can travel to the GPU box, where /root/reference does not exist.

PARITY PINNED: every function here is checked against the fixtures in
tests/golden/ which were produced by the *unmodified* reference (imported
through oracle/ref_shim.py; generator: oracle/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product package never does.

Citations are file:line into /root/reference/extended_stencil/.
"""
import math

import numpy as np

# Field order of the full coefficient sets: stencil.py:209 (2D), stencil.py:275 (3D).
FIELDS_2D = ("dt", "alphax", "alphay", "betaxy", "betayx", "deltax", "deltay")
FIELDS_3D = ("dt", "alphax", "alphay", "alphaz",
             "betaxy", "betaxz", "betayx", "betayz", "betazx", "betazy",
             "deltax", "deltay", "deltaz")

# Each stencil family = (dim, [names removed from the free parameters], [ordered copies dst<-src]).
# The copies are the `_fill_coefficients` chains flattened in MRO order (subclass first):
#   2D: stencil.py:215-218, 229-232, 243-245, 256-258, MRO note 261-266
#   3D: stencil.py:281-285, 296-303, 314-318, 328-333, 353-358, MRO notes 336-343, 361-368
_DIVFREE2 = [("betaxy", "deltay"), ("betayx", "deltax")]
_SYM2 = [("betayx", "betaxy")]
_ISO2 = [("deltay", "deltax")]
_DIVFREE3 = [("betaxy", "deltay"), ("betaxz", "deltaz"), ("betayx", "deltax"),
             ("betayz", "deltaz"), ("betazx", "deltax"), ("betazy", "deltay")]
_SYM3 = [("betayx", "betaxy"), ("betazx", "betaxz"), ("betazy", "betayz")]
_CYL3 = [("deltay", "deltax"), ("betayx", "betaxy"), ("betayz", "betaxz"), ("betazy", "betazx")]
_ISO3 = [("deltay", "deltax"), ("deltaz", "deltax"), ("betaxz", "betaxy"), ("betayz", "betaxy")]

STENCILS = {
    "StencilFree2D": (2, [], []),
    "StencilDivFree2D": (2, ["betaxy", "betayx"], _DIVFREE2),
    "StencilSymmetric2D": (2, ["betayx"], _SYM2),
    "StencilIsotropic2D": (2, ["betayx", "deltay"], _ISO2 + _SYM2),
    "StencilIsotropicDivFree2D": (2, ["betayx", "deltay", "betaxy"], _ISO2 + _SYM2 + _DIVFREE2),
    "StencilFree3D": (3, [], []),
    "StencilDivFree3D": (3, ["betaxy", "betaxz", "betayx", "betayz", "betazx", "betazy"], _DIVFREE3),
    "StencilSymmetric3D": (3, ["betayx", "betazx", "betazy"], _SYM3),
    "StencilCylinder3D": (3, ["betayx", "betayz", "betazy", "deltay"], _CYL3),
    "StencilCylinderDivFree3D": (3, ["betayx", "betayz", "betazy", "deltay",
                                     "betaxy", "betaxz", "betazx"], _CYL3 + _DIVFREE3),
    "StencilIsotropic3D": (3, ["betayx", "betazx", "betazy", "betaxz", "betayz", "deltay", "deltaz"],
                           _ISO3 + _SYM3),
    "StencilIsotropicDivFree3D": (3, ["betayx", "betazx", "betazy", "betaxz", "betayz", "deltay",
                                      "deltaz", "betaxy"], _ISO3 + _SYM3 + _DIVFREE3),
}


def fields(dim):
    return FIELDS_2D if dim == 2 else FIELDS_3D


def parameter_names(stencil):
    """Free parameters, in coefficient-field order with the fixed ones removed (stencil.py:106-114)."""
    dim, fixed, _ = STENCILS[stencil]
    fixed = set(fixed) | {f for f in fields(dim) if f.startswith("alpha")}
    return [f for f in fields(dim) if f not in fixed]


def coefficients(stencil, params):
    """Free parameters -> full coefficient vector (stencil.py:121-129 + the fill chains)."""
    dim, _, copies = STENCILS[stencil]
    names = parameter_names(stencil)
    params = np.asarray(params, dtype=float)
    if params.shape != (len(names),):
        raise ValueError("%s takes %d parameters %s" % (stencil, len(names), names))
    c = dict.fromkeys(fields(dim), 0.0)
    for n, v in zip(names, params):
        c[n] = float(v)
    for dst, src in copies:
        c[dst] = c[src]
    if dim == 2:  # stencil.py:216-217
        c["alphax"] = 1.0 - 2.0 * c["betaxy"] - 3.0 * c["deltax"]
        c["alphay"] = 1.0 - 2.0 * c["betayx"] - 3.0 * c["deltay"]
    else:         # stencil.py:282-284
        c["alphax"] = 1.0 - 2.0 * c["betaxy"] - 2.0 * c["betaxz"] - 3.0 * c["deltax"]
        c["alphay"] = 1.0 - 2.0 * c["betayx"] - 2.0 * c["betayz"] - 3.0 * c["deltay"]
        c["alphaz"] = 1.0 - 2.0 * c["betazx"] - 2.0 * c["betazy"] - 3.0 * c["deltaz"]
    return np.array([c[f] for f in fields(dim)], dtype=float)


class Grid:
    """The per-axis tables and dense k of one Dispersion object (dispersion.py:268-281, 339-357).

    `xslab=(x0, x1[, stride])` keeps only x-planes range(x0, x1, stride) of the first axis: the broadcasting
    formulae below work on any table length, which is how grids too large for one
    NumPy array are evaluated slab by slab.
    """

    def __init__(self, dim, N, Y=1.0, Z=1.0, xslab=None):
        self.dim, self.N, self.Y, self.Z = dim, N, Y, (Z if dim == 3 else 1.0)
        x = np.linspace(0, np.pi, N)
        xs = x if xslab is None else x[slice(*xslab)]     # (x0, x1) or (x0, x1, stride)
        if dim == 2:
            self.kappa = (xs.reshape(-1, 1), x.reshape(1, -1))
            self.kcomp = (self.kappa[0], self.kappa[1] / Y)
            self.k = np.sqrt(self.kcomp[0] ** 2 + self.kcomp[1] ** 2)
        else:
            self.kappa = (xs.reshape(-1, 1, 1), x.reshape(1, -1, 1), x.reshape(1, 1, -1))
            self.kcomp = (self.kappa[0], self.kappa[1] / Y, self.kappa[2] / Z)
            self.k = np.sqrt((self.kcomp[0]) ** 2 + (self.kcomp[1]) ** 2 + (self.kcomp[2]) ** 2)
        self.cos = tuple(np.cos(kap) for kap in self.kappa)
        self.s2 = tuple(0.5 * (1. - c) for c in self.cos)
        self.shape = self.k.shape

    @classmethod
    def from_tables(cls, dim, cos, s2, kcomp, Y=1.0, Z=1.0):
        """A grid built from caller-supplied per-axis tables (what the C-ABI receives)."""
        g = cls.__new__(cls)
        g.dim, g.Y, g.Z = dim, Y, (Z if dim == 3 else 1.0)
        shp = lambda a, t: np.asarray(t, dtype=float).reshape([-1 if i == a else 1 for i in range(dim)])
        g.cos = tuple(shp(a, cos[a]) for a in range(dim))
        g.s2 = tuple(shp(a, s2[a]) for a in range(dim))
        g.kcomp = tuple(shp(a, kcomp[a]) for a in range(dim))
        g.kappa = None
        g.N = g.cos[-1].size
        acc = g.kcomp[0] ** 2
        for kc in g.kcomp[1:]:
            acc = acc + kc ** 2
        g.k = np.sqrt(acc)
        g.shape = g.k.shape
        return g

    def flat_tables(self):
        """cos / s2 / k-component tables as 1-D arrays per axis (the C-ABI's table inputs)."""
        return ([c.ravel().copy() for c in self.cos], [s.ravel().copy() for s in self.s2],
                [k.ravel().copy() for k in self.kcomp])


def sqrtarg(grid, c):
    """dispersion.py:309-320 (2D) and :387-399 (3D); `c` is the full coefficient vector."""
    dx = 1.0
    if grid.dim == 2:
        dt, ax, ay, bxy, byx, dlx, dly = c
        cx, cy = grid.cos
        Ax = ax + dlx * (1.0 + 2.0 * cx) + 2. * bxy * cy
        Ay = ay + dly * (1.0 + 2.0 * cy) + 2. * byx * cx
        return Ax * grid.s2[0] / (dx ** 2) + Ay * grid.s2[1] / ((grid.Y * dx) ** 2)
    dt, ax, ay, az, bxy, bxz, byx, byz, bzx, bzy, dlx, dly, dlz = c
    cx, cy, cz = grid.cos
    Ax = ax + dlx * (1.0 + 2.0 * cx) + 2. * bxy * cy + 2. * bxz * cz
    Ay = ay + dly * (1.0 + 2.0 * cy) + 2. * byx * cx + 2. * byz * cz
    Az = az + dlz * (1.0 + 2.0 * cz) + 2. * bzx * cx + 2. * bzy * cy
    return (Ax * grid.s2[0] / (dx ** 2) + Ay * grid.s2[1] / ((grid.Y * dx) ** 2)
            + Az * grid.s2[2] / ((grid.Z * dx) ** 2))


def corner_values(dim, c, Y=1.0, Z=1.0):
    """Closed-form sqrtarg at the non-zero corners kappa in {0,pi}^dim (dispersion.py:294-297, 369-375)."""
    if dim == 2:
        dt, ax, ay, bxy, byx, dlx, dly = c
        return [(ay + 2 * byx - dly) / (Y ** 2),
                ax - 2 * bxy - dlx + (ay - 2 * byx - dly) / (Y ** 2),
                ax + 2 * bxy - dlx]
    dt, ax, ay, az, bxy, bxz, byx, byz, bzx, bzy, dlx, dly, dlz = c
    return [(az + 2.0 * bzx + 2.0 * bzy - dlz) / Z ** 2,
            (ay + 2.0 * byx + 2.0 * byz - dly) / Y ** 2,
            (ay + 2.0 * byx - 2.0 * byz - dly) / Y ** 2 + (az + 2.0 * bzx - 2.0 * bzy - dlz) / Z ** 2,
            ax + 2.0 * bxy + 2.0 * bxz - dlx,
            ax + 2.0 * bxy - 2.0 * bxz - dlx + (az - 2.0 * bzx + 2.0 * bzy - dlz) / Z ** 2,
            ax - 2.0 * bxy + 2.0 * bxz - dlx + (ay - 2.0 * byx + 2.0 * byz - dly) / Y ** 2,
            ax - 2.0 * bxy - 2.0 * bxz - dlx + (ay - 2.0 * byx - 2.0 * byz - dly) / Y ** 2
            + (az - 2.0 * bzx - 2.0 * bzy - dlz) / Z ** 2]


def weight(grid, kind="equal", params=()):
    """weight.py:4-31; 'equal' is the class default w = 1.0 (dispersion.py:40)."""
    if kind == "equal":
        return 1.0
    if kind == "xaxis":
        w = np.zeros_like(grid.k)
        if grid.dim == 2:
            w[:, 0] = 1.0
        else:
            w[:, 0, 0] = 1.0
        return w
    if kind == "xaxis_soft":
        width = params[0] if len(params) else 0.1
        if grid.dim == 2:
            return np.exp(-(grid.kcomp[1] / width) ** 2)
        return np.exp(-(grid.kcomp[1] / width) ** 2 - (grid.kcomp[2] / width) ** 2)
    raise ValueError(kind)


def raw_parts(grid, c):
    """(Amin, Amax) of sqrtarg and the array itself -- the quantities the device kernel reduces."""
    A = sqrtarg(grid, c)
    return A, np.min(A), np.max(A)


def gate(dim, c, Amin, Amax, Y=1.0, Z=1.0, dt_multiplier=1.0):
    """stencil_ok / dt_ok from the grid extrema (dispersion.py:284-307, 360-385, 129-154).

    np.max(dx*sqrt(A)) == sqrt(max A): IEEE sqrt is monotone and dx == 1.
    Returns (stencil_ok, dt_ok, passed) with passed == "omega() would not return None".
    """
    if np.isnan(Amin) or not (Amin < 0):
        sok = min(corner_values(dim, c, Y, Z))
    else:
        sok = Amin
    if math.isnan(sok):
        raise ValueError("stencil_ok got NaN")
    if sok < 0:
        return sok, sok, False
    with np.errstate(divide="ignore", invalid="ignore"):
        dtok = dt_multiplier / np.sqrt(Amax) - c[0]
    if math.isnan(dtok):
        raise ValueError("dt_ok got NaN")
    return sok, dtok, not (dtok < 0 or c[0] <= 0)


def omega_from_sqrtarg(A, dt):
    """dispersion.py:111-121 and :186-191 (dx = 1)."""
    dx = 1.0
    sqrtres = np.sqrt(A)
    return (2. / (dt * dx)) * np.arcsin(dt * dx * sqrtres)


def norm_scale(dim, N, Y=1.0, Z=1.0):
    """dispersion.py:253 -- the factor applied to the raw weighted sum."""
    return (np.pi / (N - 1)) ** dim * Y ** 2 * Z ** 2


def evaluate(grid, c, w=1.0, dt_multiplier=1.0):
    """One objective evaluation for full coefficients `c`.

    Returns a dict with the device-contract quantities (S, Amin, Amax, nan) and the
    reference-API quantities (stencil_ok, dt_ok, omega or None, norm).
    """
    c = np.asarray(c, dtype=float)
    out = {}
    if np.any(np.isnan(c)):  # dispersion.py:241-243, 137-139, 286-288
        out.update(S=np.nan, Amin=np.nan, Amax=np.nan, nan=0, stencil_ok=-1, dt_ok=-1., omega=None, norm=1e6)
        return out
    with np.errstate(invalid="ignore"):
        A, Amin, Amax = raw_parts(grid, c)
        sok, dtok, passed = gate(grid.dim, c, Amin, Amax, grid.Y, grid.Z, dt_multiplier)
        out.update(Amin=Amin, Amax=Amax, stencil_ok=sok, dt_ok=dtok)
        if not passed:
            out.update(S=np.nan, nan=0, omega=None, norm=1e6)  # dispersion.py:246-248
            return out
        om = omega_from_sqrtarg(A, c[0])
    nan = int(np.count_nonzero(np.isnan(om)))
    if nan:
        raise ValueError("Encountered NaNs in omega for coefficients", c)  # dispersion.py:193-194
    f = w * (om - grid.k) ** 2          # dispersion.py:251
    S = f.sum()                          # dispersion.py:252
    out.update(S=S, nan=nan, omega=om, norm=S * norm_scale(grid.dim, grid.N, grid.Y, grid.Z))
    return out


def device_contract(grid, c, w=1.0):
    """Un-gated (S, Amin, Amax, nan) exactly as the CUDA kernel is specified to return them:
    S = sum w*(omega-k)^2 over the (slab of the) grid computed for EVERY candidate, gate or not
    (NaN where sqrtarg < 0 or dt*sqrtres > 1)."""
    c = np.asarray(c, dtype=float)
    with np.errstate(invalid="ignore", divide="ignore"):
        A, Amin, Amax = raw_parts(grid, c)
        om = omega_from_sqrtarg(A, c[0])
        S = (w * (om - grid.k) ** 2).sum()
    return S, Amin, Amax, int(np.count_nonzero(np.isnan(om)))


def norm(stencil, params, dim, N, Y=1.0, Z=1.0, wkind="equal", wparams=(), dt_multiplier=1.0, grid=None):
    """Convenience: `Dispersion{2,3}D(...).norm(params)` for a named stencil family."""
    grid = grid or Grid(dim, N, Y, Z)
    if np.any(np.isnan(np.asarray(params, dtype=float))):
        return 1e6
    return evaluate(grid, coefficients(stencil, params), weight(grid, wkind, wparams), dt_multiplier)["norm"]
