"""Stencil families: free parameters -> full coefficient sets (drop-in for the reference's
extended_stencil/stencil.py; same class names, flags, `Parameters` / `Coefficients` / `dof` /
`coefficients()` / `get_stencil()` surface).

Written table-first instead of as chained `_fill_coefficients` overrides: every family declares the
ordered list of "tie" assignments (dst <- src) that the reference performs in MRO order
(stencil.py:215-218, 229-232, 243-245, 256-258, 281-285, 296-303, 314-318, 328-333, 353-358 and
the MRO notes at :261-266, :336-343, :361-368).  From that table the base class derives

  * the free-parameter list (stencil.py:106-114),
  * a scalar fill returning the same shape-(1,) record array the reference returns, and
  * `coefficients_batch(P)`, a vectorised fill of a whole (B, dof) parameter matrix -- what the
    batched GPU objective consumes (the reference cannot do this: SURVEY.md section 7, "Batching
    without changing the search").

Host-side and O(13) flops per vector, so it stays NumPy; the alpha formulae keep the reference's
operation order so that the coefficients are bit-identical.
"""
from enum import Enum

import numpy as np


class namedarray:
    """A list of field names that can cast number sequences into a one-row record array."""

    def __init__(self, field_names):
        if isinstance(field_names, str):
            field_names = [f.strip() for f in field_names.split(',')]
        self._fields = list(field_names)
        self._nfields = len(self._fields)
        self.dtype = np.dtype([(f, float) for f in self._fields])

    def __repr__(self):
        return '{}({})'.format(type(self).__name__, self._fields)

    def __call__(self, args):
        return np.asarray(args, dtype=float).view(dtype=self.dtype).view(np.recarray)

    def convert(self, other):
        """Widen a record array to this field set; fields it does not have become 0."""
        wide = self(np.zeros(self._nfields))
        for f in other.dtype.names:
            wide[f] = other[f]
        return wide


class StencilFlags(Enum):
    DIM2 = 0
    DIM3 = 1
    DIVFREE = 2
    ISOTROPIC = 3
    CYLINDERSYM = 4
    SYMMETRICBETA = 5


_F = StencilFlags

_COEFF_2D = "dt, alphax, alphay, betaxy, betayx, deltax, deltay"
_COEFF_3D = ("dt, alphax, alphay, alphaz, betaxy, betaxz, betayx, betayz, betazx, betazy, "
             "deltax, deltay, deltaz")

# alpha_i = 1 - 2*beta_ij [- 2*beta_ik] - 3*delta_i, left to right  (stencil.py:216-217, 282-284)
_ALPHA_2D = (("alphax", ("betaxy",), "deltax"), ("alphay", ("betayx",), "deltay"))
_ALPHA_3D = (("alphax", ("betaxy", "betaxz"), "deltax"), ("alphay", ("betayx", "betayz"), "deltay"),
             ("alphaz", ("betazx", "betazy"), "deltaz"))


class Stencil:
    """Base of all stencil families."""

    Coefficients = None          # namedarray of the full coefficient set
    Parameters = None            # namedarray of the free parameters (per instance)
    Flags = frozenset()
    TIES = ()                    # ordered (dst, src) copies, already in the reference's MRO order
    ALPHAS = ()
    stencils = {}                # flag set -> class, filled by register_subclasses()
    _order = {}                  # class -> registration rank (tie-break of get_stencil)

    def __init__(self):
        tied = {dst for dst, _ in self.TIES} | {a[0] for a in self.ALPHAS}
        self._parameters = [f for f in self.Coefficients._fields if f not in tied]
        self.Parameters = namedarray(self._parameters)
        col = {f: i for i, f in enumerate(self.Coefficients._fields)}
        self._param_cols = np.array([col[f] for f in self._parameters])
        self._tie_cols = [(col[d], col[s]) for d, s in self.TIES]
        self._alpha_cols = [(col[a], [col[b] for b in betas], col[d]) for a, betas, d in self.ALPHAS]
        # The same procedure "compiled" into a handful of array operations (coefficients_batch runs once per
        # SLSQP iterate): the ordered copies collapse into one gather from the parameter columns, and the
        # alphas -- which only read betas and deltas -- are evaluated side by side, element by element in the
        # reference's own order (1 - 2 b1 [- 2 b2] - 3 d), so every coefficient keeps its bits.
        src = [-1] * self.Coefficients._nfields
        for k, c in enumerate(self._param_cols):
            src[c] = k
        for d, s_ in self._tie_cols:
            src[d] = src[s_]
        self._gather_dst = np.array([c for c in range(len(src)) if src[c] >= 0], dtype=int)
        self._gather_src = np.array([src[c] for c in self._gather_dst], dtype=int)
        acols = [a for a, _, _ in self._alpha_cols]
        nb = {len(b) for _, b, _ in self._alpha_cols}
        reads = {c for _, b, d in self._alpha_cols for c in list(b) + [d]}
        self._alpha_plan = None
        if len(nb) == 1 and not (reads & set(acols)):
            self._alpha_plan = (np.array(acols, dtype=int),
                                [np.array([b[k] for _, b, _ in self._alpha_cols], dtype=int) for k in range(nb.pop())],
                                np.array([d for _, _, d in self._alpha_cols], dtype=int))

    @property
    def dof(self):
        return len(self._parameters)

    @property
    def ncoef(self):
        return self.Coefficients._nfields

    def coefficients_batch(self, params):
        """(B, dof) free parameters -> (B, ncoef) full coefficients (float64, C order)."""
        P = np.asarray(params, dtype=float)
        if P.ndim == 1:
            P = P.reshape(1, -1)
        if P.ndim != 2 or P.shape[1] != self.dof:
            raise ValueError('{} takes {} parameters ({}), got shape {}'.format(
                type(self).__name__, self.dof, ', '.join(self._parameters), P.shape))
        C = np.zeros((P.shape[0], self.ncoef))
        C[:, self._gather_dst] = P[:, self._gather_src]
        if self._alpha_plan is not None:
            acols, bcols, dcols = self._alpha_plan
            acc = 1.0 - 2.0 * C[:, bcols[0]]
            for b in bcols[1:]:
                acc = acc - 2.0 * C[:, b]
            C[:, acols] = acc - 3.0 * C[:, dcols]
        else:
            for a, betas, d in self._alpha_cols:
                acc = 1.0 - 2.0 * C[:, betas[0]]
                for b in betas[1:]:
                    acc = acc - 2.0 * C[:, b]
                C[:, a] = acc - 3.0 * C[:, d]
        return C

    def _coefficients_batch_sequential(self, params):
        """The uncompiled procedure (ordered copies, then one alpha after the other): the check of the above."""
        P = np.asarray(params, dtype=float).reshape(-1, self.dof)
        C = np.zeros((P.shape[0], self.ncoef))
        C[:, self._param_cols] = P
        for dst, src in self._tie_cols:
            C[:, dst] = C[:, src]
        for a, betas, d in self._alpha_cols:
            acc = 1.0 - 2.0 * C[:, betas[0]]
            for b in betas[1:]:
                acc = acc - 2.0 * C[:, b]
            C[:, a] = acc - 3.0 * C[:, d]
        return C

    def coefficients(self, args):
        """Free parameters -> one-row record array of all coefficients (as the reference returns)."""
        P = np.asarray(args, dtype=float).ravel()
        return self.Coefficients(self.coefficients_batch(P)[0])

    # -- registry / selection (stencil.py:138-199) -------------------------------------------------
    @classmethod
    def register_subclasses(cls, sub=None):
        for subclass in (cls if sub is None else sub).__subclasses__():
            Stencil.register_subclasses(subclass)
            Stencil.stencils[subclass.Flags] = subclass
            Stencil._order.setdefault(subclass, len(Stencil._order))

    @classmethod
    def get_stencil(cls, args):
        """Smallest registered flag set that contains what `args` asks for."""
        if not Stencil.stencils:
            Stencil.register_subclasses()
        if args.dim not in (2, 3):
            raise NotImplementedError
        wanted = {_F.DIM2 if args.dim == 2 else _F.DIM3}
        if args.div_free:
            wanted.add(_F.DIVFREE)
        if args.symmetric_beta:
            wanted.add(_F.SYMMETRICBETA)
        if args.symmetric_axes == args.dim - 1:
            wanted.add(_F.ISOTROPIC)
        if args.symmetric_axes == 1 and args.dim == 3:
            wanted.add(_F.CYLINDERSYM)
        fits = [k for k in Stencil.stencils if k >= wanted]
        if not fits:
            raise NotImplementedError
        best = min(fits, key=lambda k: (len(k), Stencil._order[Stencil.stencils[k]]))
        if best > wanted:
            print('Chosen stencil also has flags: ', best - wanted)
        return Stencil.stencils[best]()


def get_stencil(args):
    return Stencil.get_stencil(args)


# ---- tie tables (dst <- src), named after what they enforce -----------------------------------------
_DIVFREE_2D = (("betaxy", "deltay"), ("betayx", "deltax"))
_SYMBETA_2D = (("betayx", "betaxy"),)
_ISO_2D = (("deltay", "deltax"),)
_DIVFREE_3D = (("betaxy", "deltay"), ("betaxz", "deltaz"), ("betayx", "deltax"),
               ("betayz", "deltaz"), ("betazx", "deltax"), ("betazy", "deltay"))
_SYMBETA_3D = (("betayx", "betaxy"), ("betazx", "betaxz"), ("betazy", "betayz"))
_CYL_3D = (("deltay", "deltax"), ("betayx", "betaxy"), ("betayz", "betaxz"), ("betazy", "betazx"))
_ISO_3D = (("deltay", "deltax"), ("deltaz", "deltax"), ("betaxz", "betaxy"), ("betayz", "betaxy"))


class StencilFree2D(Stencil):
    """Most general two-dimensional stencil."""
    Coefficients = namedarray(_COEFF_2D)
    Flags = frozenset([_F.DIM2])
    ALPHAS = _ALPHA_2D


class StencilDivFree2D(StencilFree2D):
    """2D, preserves div B = 0."""
    Flags = StencilFree2D.Flags | {_F.DIVFREE}
    TIES = _DIVFREE_2D


class StencilSymmetric2D(StencilFree2D):
    """2D, symmetric beta_ij."""
    Flags = StencilFree2D.Flags | {_F.SYMMETRICBETA}
    TIES = _SYMBETA_2D


class StencilIsotropic2D(StencilSymmetric2D):
    """2D, identical on both axes."""
    Flags = StencilSymmetric2D.Flags | {_F.ISOTROPIC}
    TIES = _ISO_2D + _SYMBETA_2D


class StencilIsotropicDivFree2D(StencilIsotropic2D, StencilDivFree2D):
    """2D, identical on both axes and div-free: deltas are tied first, then betas follow the deltas."""
    Flags = StencilDivFree2D.Flags | StencilIsotropic2D.Flags
    TIES = _ISO_2D + _SYMBETA_2D + _DIVFREE_2D


class StencilFree3D(Stencil):
    """Most general three-dimensional stencil."""
    Coefficients = namedarray(_COEFF_3D)
    Flags = frozenset([_F.DIM3])
    ALPHAS = _ALPHA_3D


class StencilDivFree3D(StencilFree3D):
    """3D, preserves div B = 0."""
    Flags = StencilFree3D.Flags | {_F.DIVFREE}
    TIES = _DIVFREE_3D


class StencilSymmetric3D(StencilFree3D):
    """3D, symmetric beta_ij."""
    Flags = StencilFree3D.Flags | {_F.SYMMETRICBETA}
    TIES = _SYMBETA_3D


class StencilCylinder3D(StencilFree3D):
    """3D, two of the three axes identical."""
    Flags = StencilFree3D.Flags | {_F.CYLINDERSYM}
    TIES = _CYL_3D


class StencilCylinderDivFree3D(StencilCylinder3D, StencilDivFree3D):
    """3D, two identical axes and div-free."""
    Flags = StencilCylinder3D.Flags | StencilDivFree3D.Flags
    TIES = _CYL_3D + _DIVFREE_3D


class StencilIsotropic3D(StencilSymmetric3D):
    """3D, symmetric beta_ij, identical on all axes."""
    Flags = StencilSymmetric3D.Flags | {_F.ISOTROPIC}
    TIES = _ISO_3D + _SYMBETA_3D


class StencilIsotropicDivFree3D(StencilIsotropic3D, StencilDivFree3D):
    """3D, identical on all axes and div-free."""
    Flags = StencilIsotropic3D.Flags | StencilDivFree3D.Flags
    TIES = _ISO_3D + _SYMBETA_3D + _DIVFREE_3D
