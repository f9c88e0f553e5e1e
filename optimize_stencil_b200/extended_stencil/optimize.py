"""Multi-start SLSQP driver (drop-in for the reference's extended_stencil/optimize.py).

SciPy still drives every search (`scipy.optimize.minimize(method='SLSQP')` with the reference's
constraints and options, optimize.py:136,142).  What changes is the data-parallel part:

  reference: mp.Pool(nproc).imap(_optimize_single, starts) -- one whole SLSQP run per CPU worker,
             every objective / constraint call a fresh ~25-pass NumPy sweep (optimize.py:169-190);
  here:      the searches run in this process against the GPU-backed Dispersion objects.  The start
             grid is screened in TWO batched launches (CFL time step of every start, then
             feasibility of every start), and inside each search one launch serves the objective,
             both constraints and all their finite-difference points of an iterate (see
             dispersion.py).  `--singlecore` is accepted and changes nothing by default.
             SLSQP is handed the forward differences it would take itself as `jac` callables and the
             constraint list stacked into one vector-valued constraint (`exact_fd_jac`: same points, same
             arithmetic, same rows -- the search is bit-identical, the Python overhead per iterate drops 3x).
             What remains is SciPy's own per-iteration Python time.  It can be spread over worker processes
             that share the GPU (`args.workers = N` or the environment variable OPTSTENCIL_WORKERS=N; every
             worker re-creates its device context lazily, exactly like the reference's pickled Dispersion
             objects re-run init2; each start is independent and deterministic, so the result does not depend
             on the number of workers) -- but measured on one B200 without MPS this does not pay: the
             workers' small launches time-slice the device (4096 starts on 128^3: 29.7 s in-process, 31.5 s
             with 8 workers, tools/workers_timing.py), so it stays opt-in.

The printed lines keep the reference's format (optimize.py:71, 144, 191, 200-202).
"""
import itertools
import os
import sys

import numpy as np
import scipy.optimize as scop

from .dispersion import Dispersion2D, Dispersion3D, FD_STEP
from .stencil import get_stencil
from .weight import weight_functions


class _single_constraint:
    """Box constraint on one parameter.  `jac` is SciPy's own forward difference of it
    (approx_derivative, '2-point', abs_step=FD_STEP: (f(x + h e_i) - f(x)) / ((x_i + h) - x_i)), evaluated
    in closed form; it is what SLSQP would compute itself, minus the Python overhead."""

    def jac(self, p):
        p = np.asarray(p, dtype=float)
        i = self.index
        xi = p[i]
        dx = (xi + FD_STEP) - xi
        f0 = self(p)
        if dx == 0 or not np.isfinite(f0):
            from scipy.optimize._numdiff import approx_derivative
            return approx_derivative(self, p, method='2-point', abs_step=FD_STEP)
        J = np.zeros(len(p))
        x1 = p.copy()
        x1[i] = xi + FD_STEP
        J[i] = (self(x1) - f0) / dx
        return J


class make_single_max_constraint(_single_constraint):
    """p[index] <= maxval as an SLSQP 'ineq' function."""

    def __init__(self, index, maxval):
        self.index, self.maxval = index, maxval

    def __call__(self, p):
        return self.maxval - p[self.index]


class make_single_min_constraint(_single_constraint):
    """p[index] >= minval as an SLSQP 'ineq' function."""

    def __init__(self, index, minval):
        self.index, self.minval = index, minval

    def __call__(self, p):
        return p[self.index] - self.minval


class _stacked_constraints:
    """The reference's list of scalar 'ineq' constraints (optimize.py:96-110) as ONE vector-valued SLSQP
    constraint with the rows in the same order -- SLSQP sees the same constraint vector and the same normals,
    bit for bit, but Python is entered once instead of 2 + 2*dof times per evaluation."""

    def __init__(self, dispersion, cons):
        self.dispersion = dispersion
        self.box = [c['fun'] for c in cons[2:]]
        self.index = np.array([b.index for b in self.box], dtype=int)
        self.is_min = np.array([isinstance(b, make_single_min_constraint) for b in self.box])
        self.bound = np.array([b.minval if m else b.maxval for b, m in zip(self.box, self.is_min)], dtype=float)
        self.rows = np.arange(len(self.box))

    def _box_values(self, x):
        return np.where(self.is_min, x - self.bound, self.bound - x)      # p[i] - minval | maxval - p[i]

    def fun(self, p):
        d = self.dispersion
        p = np.asarray(p, dtype=float)
        out = np.empty(2 + len(self.box))
        out[0] = np.asarray(d.stencil_ok(p), dtype=float).ravel()[0]
        out[1] = np.asarray(d.dt_ok(p), dtype=float).ravel()[0]
        out[2:] = self._box_values(p[self.index])
        return out

    def jac(self, p):
        d = self.dispersion
        p = np.asarray(p, dtype=float)
        J = np.zeros((2 + len(self.box), len(p)))
        J[0] = d.stencil_ok_fd(p)
        J[1] = d.dt_ok_fd(p)
        xi = p[self.index]
        xh = xi + FD_STEP
        dx = xh - xi
        f0 = self._box_values(xi)
        if np.all(dx) and np.all(np.isfinite(f0)):
            J[2 + self.rows, self.index] = (self._box_values(xh) - f0) / dx   # scipy's (f(x + h e_i) - f(x)) / dx
        else:
            for r, b in enumerate(self.box):
                J[2 + r] = b.jac(p)
        return J


SLSQP_OPTIONS = dict(disp=False, iprint=2)


class Optimize:
    """Holds the low- and high-resolution dispersion objects and runs the start-grid scan."""

    # Hand SciPy the finite differences it would take anyway (same points, same arithmetic, from the
    # result cache).  The search is bit-identical either way; OPTSTENCIL_FD_JAC=0 lets SciPy do them.
    exact_fd_jac = os.environ.get('OPTSTENCIL_FD_JAC', '1') != '0'

    def __init__(self, args):
        self.args = args
        self.dim = args.dim
        self.div_free = args.div_free
        self.stencil = get_stencil(args)
        print('Using stencil', self.stencil.__class__.__name__, 'with free parameters',
              ", ".join(self.stencil.Parameters._fields))

        def make(N):
            if self.dim == 2:
                return Dispersion2D(args.Y, N, self.stencil)
            return Dispersion3D(args.Y, args.Z, N, self.stencil)

        self.dispersion = make(args.Ngrid_low)
        self.dispersion_high = self.dispersion if args.Ngrid_high == args.Ngrid_low else make(args.Ngrid_high)
        for d in {id(self.dispersion): self.dispersion, id(self.dispersion_high): self.dispersion_high}.values():
            d.dt_multiplier = args.dt_multiplier
            if args.weight in weight_functions:
                weight_functions[args.weight](*args.weight_params)(d)

        self.dtmin, self.dtmax = self.args.dtrange[0], self.args.dtrange[1]
        self.constraints = self._constraint_list(self.dispersion)
        self.constraints_high = self._constraint_list(self.dispersion_high)

    def _constraint_list(self, dispersion):
        cons = [dict(type='ineq', fun=dispersion.stencil_ok), dict(type='ineq', fun=dispersion.dt_ok)]
        for i, fname in enumerate(self.stencil.Parameters._fields):
            lo, hi = getattr(self.args, fname + 'range')
            cons.append(dict(type='ineq', fun=make_single_min_constraint(i, lo)))
            cons.append(dict(type='ineq', fun=make_single_max_constraint(i, hi)))
        if self.exact_fd_jac:
            cons[0]['jac'], cons[1]['jac'] = dispersion.stencil_ok_fd, dispersion.dt_ok_fd
            for con in cons[2:]:
                con['jac'] = con['fun'].jac
        return cons

    def _minimize(self, dispersion, constraints, x0):
        """scipy.optimize.minimize exactly as the reference calls it (optimize.py:136, 142).  With
        `exact_fd_jac` the forward differences SLSQP would take itself are supplied as `jac`, and the
        constraint list is handed over stacked into one vector-valued constraint (same rows, same order)."""
        if not self.exact_fd_jac:
            return scop.minimize(dispersion.norm, x0, method='SLSQP', constraints=constraints,
                                 options=dict(SLSQP_OPTIONS))
        stacked = _stacked_constraints(dispersion, constraints)
        return scop.minimize(dispersion.norm, x0, method='SLSQP', jac=dispersion.norm_fd,
                             constraints=[dict(type='ineq', fun=stacked.fun, jac=stacked.jac)],
                             options=dict(SLSQP_OPTIONS))

    # -- one start (optimize.py:115-139) ----------------------------------------------------------
    def _optimize_single(self, x0):
        x0 = list(x0)
        if x0[0] is None:
            x0[0] = 0
            dt_ok = np.asarray(self.dispersion.dt_ok(x0)).item()
            if dt_ok < 0:
                return x0, None, float('inf')
            x0[0] = max(min(dt_ok, self.dtmax), self.dtmin)
        x0 = np.asarray(x0, dtype=float)
        if self.dispersion.stencil_ok(x0) < 0:
            return x0, None, float('inf')
        res = self._minimize(self.dispersion, self.constraints, x0)
        return x0, res, self.dispersion_high.norm(res.x)

    def _optimize_final(self, x0):
        res = self._minimize(self.dispersion_high, self.constraints_high, x0)
        print('x0={}, x={}, norm={}'.format(x0, res.x, res.fun))
        return res, res.fun

    # -- the scan (optimize.py:147-205) -----------------------------------------------------------
    def start_grid(self):
        """Interior points of an (N+2)-point linspace per parameter; dt scanned or left to the CFL rule."""
        N = self.args.N
        axes = []
        for fname in self.stencil.Parameters._fields[1:]:
            lo, hi = getattr(self.args, fname + 'range')
            axes.append(np.linspace(lo, hi, N + 2)[1:-1])
        dts = np.linspace(self.dtmin, self.dtmax, N + 2)[1:-1] if self.args.scan_dt else [None]
        return list(itertools.product(dts, *axes))

    def _prescreen(self, starts):
        """Warm the result cache for the whole start grid with batched launches (replaces the per-start
        NumPy sweeps the pool workers did first: optimize.py:118-134)."""
        rows = [[0.0 if v is None else v for v in s] for s in starts]
        if not rows:
            return
        P = np.asarray(rows, dtype=float)
        first = self.dispersion.evaluate_batch(P)
        if starts[0][0] is None:   # dt comes from the CFL limit at dt = 0, clipped to dtrange
            ok = first['dt_ok'] >= 0
            P = P[ok]
            P[:, 0] = np.clip(first['dt_ok'][ok], self.dtmin, self.dtmax)
            if len(P):
                self.dispersion.evaluate_batch(P)

    def optimize(self):
        """Scan the grid of starting values, keep the best local optimum, polish it at high resolution."""
        starts = self.start_grid()
        workers = int(getattr(self.args, 'workers', 0) or os.environ.get('OPTSTENCIL_WORKERS', 0) or 0)
        pool = None
        method = os.environ.get('OPTSTENCIL_START_METHOD', 'spawn')       # never fork a live CUDA context
        if workers > 1 and method == 'spawn' and not hasattr(sys.modules.get('__main__'), '__file__'):
            print('optimize: worker processes need an importable __main__ (not stdin / REPL); running in-process')
            workers = 0
        if workers > 1 and len(starts) >= 2 * workers and not getattr(self.args, 'singlecore', False):
            import multiprocessing as mp
            pool = mp.get_context(method).Pool(workers)
            results = pool.imap(self._optimize_single, starts, chunksize=max(1, len(starts) // (8 * workers)))
        else:
            self._prescreen(starts)
            results = map(self._optimize_single, starts)
        best = None
        for x0, res, norm in results:
            print('x0={}, x={}, norm={}'.format(x0, res.x if res else None, norm))
            sys.stdout.flush()
            if best is None or norm < best[2]:
                best = (x0, res, norm)
        if pool is not None:
            pool.close()
            pool.join()
        bestx0, bestres, bestnorm = best
        if self.args.Ngrid_high != self.args.Ngrid_low:
            print()
            print('using this result as starting point for high-res optimization:')
            print('x0={},x={}, norm={}'.format(bestx0, bestres.x, bestnorm))
            bestres, bestnorm = self._optimize_final(bestres.x)
        return bestres.x, bestnorm

    def omega_output(self, fname, x):
        return self.dispersion_high.omega_output(fname, x)
