"""Numerical dispersion relation of a stencil on a dense k-grid -- GPU-backed drop-in for the
reference's extended_stencil/dispersion.py.

Same public surface (`Dispersion2D(Y, N, stencil)`, `Dispersion3D(Y, Z, N, stencil)`, `norm`,
`omega`, `stencil_ok`, `dt_ok`, `omega_output`, `omega_spline` / `omega_interp`, the `parameters`,
`coefficients`, `sqrtarg`, `sqrtres` properties and the table attributes), same sentinel / error
behaviour (1e6, None, -1, ValueError).  What differs is where the grid arithmetic happens:

  reference: ~25 NumPy passes over N^dim temporaries per norm() call          (dispersion.py:235-255)
  here:      one fused sm_100a kernel returning (S, Amin, Amax, nan) per coefficient vector, from
             which norm / stencil_ok / dt_ok / "is omega None" follow on the host ("gate" below).

There is NO CPU fallback: the per-axis tables are NumPy (exactly the reference's init2 formulae),
every N^dim quantity comes from optimize_stencil_b200._capi.Context, which raises when the CUDA
library or a device is missing.

Additions that do not exist in the reference:
  * evaluate_batch / norm_batch: a whole matrix of parameter vectors per launch;
  * a result cache keyed by the parameter bytes (generalises the one-entry cache at
    dispersion.py:94-96) plus speculative evaluation of the n forward-difference neighbours SciPy's
    SLSQP is about to request (same absolute step 1.4901161193847656e-08), so that the objective, both
    constraints and all their finite-difference Jacobians at one iterate cost ONE launch;
  * `k` (dense) is built lazily -- the device recomputes k per point and never needs it;
  * `devices`: batches large enough to pay for it are sharded over several GPUs of this process by the
    library's device group (sb200_group_*), replacing the reference's mp.Pool scan (optimize.py:169-190);
  * `view()`: a cheap per-search facade sharing the tables and the device context, with its own result
    cache -- what lets `Optimize` run many SLSQP searches in lock-step, one launch per round.

Determinism: every result a search sees comes from a launch (or launch SEGMENT) made of exactly the points
that search asked for -- an iterate and its forward-difference neighbours -- and a finite difference never
mixes two launches (cache entries carry their launch id).  The library guarantees that a segment's results do
not depend on what else shares the launch, so a search is bit-identical whether it runs alone or next to
thousands of others.
"""
import os

import numpy as np
import scipy.interpolate as spinterp

from .. import _capi

FD_STEP = 1.4901161193847656e-08   # scipy/optimize/_slsqp_py.py: _epsilon = sqrt(finfo(float).eps)


def _python_min(columns):
    """Element-wise equivalent of Python's min() over a list (the reference calls min(b) on its corner
    list): keep the first entry unless a later one is strictly smaller -- NaNs behave as they do there."""
    best = columns[0]
    for col in columns[1:]:
        best = np.where(col < best, col, best)
    return best


class Dispersion:
    """Base class; use Dispersion2D or Dispersion3D."""

    dx = 1.0
    Y = 1.0
    Z = 1.0
    dim = 0
    dt_multiplier = 1.0
    stencil = None
    device = None              # CUDA ordinal of a single-device context (None: $OPTSTENCIL_DEVICE, else 0)
    devices = None             # None: automatic (one device; every visible device once a batch is large enough to
                               # pay for the extra contexts, or what $OPTSTENCIL_DEVICES says: "all", "0,1,..."),
                               # 'all', or a sequence of CUDA ordinals
    throughput_evals = float(2 ** 32)   # evaluations in ONE evaluate_batch call from which it is a throughput launch (fast
                                        # series, no cache warming) even if its rows would fit the per-point cache
    auto_multi_evals = float(2 ** 36)   # k-point x candidate evaluations in ONE batch from which `devices = None`
                                        # brings up the other GPUs (~0.1 s of one B200; a context costs ~0.3 s once)
    fd_prefetch = True         # evaluate x + h*e_i alongside every cache miss (see module docstring)
    cache_size = 4096

    def __init__(self, runinit2=True):
        self._parameters = None
        self._coefficients = None
        self._w = 1.0
        self._w_dirty = True
        self._ctx = None
        self._cache = {}
        self._fd_last = None
        self._fd_index = {}
        self._launch_id = 0
        self._submit = None        # views only: hands a block of coefficient rows to the lock-step server
        self._parent = None        # views only
        self._k = None
        self.init2_run = False
        self.stats = dict(launches=0, evaluated=0, hits=0, misses=0)
        if runinit2:
            self.init2()

    # -- tables (dispersion.py:268-281, 339-357) -------------------------------------------------
    def init2(self):
        x = np.linspace(0, np.pi, self.N)
        self.kappamesh = np.meshgrid(*([x] * self.dim), indexing='ij', sparse=True)
        scale = (1.0, self.Y, self.Z)
        self.kmesh = tuple(kap if a == 0 else kap / scale[a] for a, kap in enumerate(self.kappamesh))
        cos = [np.cos(kap) for kap in self.kappamesh]
        s2 = [0.5 * (1. - c) for c in cos]
        for a, name in enumerate('xyz'[:self.dim]):
            setattr(self, 'kappa' + name, self.kappamesh[a])
            setattr(self, 'k' + name, self.kmesh[a])
            setattr(self, 'coskappa' + name, cos[a])
            setattr(self, 's' + name + '2', s2[a])
        self.dks = tuple(km.ravel()[1] for km in self.kmesh)
        self._tables = (cos, s2, list(self.kmesh))
        self.init2_run = True

    @property
    def k(self):
        """|k| on the dense grid (dispersion.py:276, 350).  Built on first use only."""
        if not self.init2_run:
            self.init2()
        if self._k is None:
            acc = self.kmesh[0] ** 2
            for km in self.kmesh[1:]:
                acc = acc + km ** 2
            self._k = np.sqrt(acc)
        return self._k

    # -- pickling: the device handle is rebuilt lazily, like the reference's init2 state ----------
    def __getstate__(self):
        state = self.__dict__.copy()
        state['_ctx'] = None
        state['_w_dirty'] = True
        state['_cache'] = {}
        state['_fd_last'] = None
        state['_submit'] = None
        state['_parent'] = None
        return state

    # -- per-search views (new) -------------------------------------------------------------------
    def view(self, submit=None):
        """A facade for ONE search: same tables, stencil, weight and device context as this object, but its own
        parameters, result cache and finite-difference memo.  `submit(C) -> (S, Amin, Amax, nan)` replaces the
        direct launch (the lock-step server of optimize.py passes its own); None launches directly."""
        if submit is None:
            self._context()                  # tables, context and weight upload happen once, on the parent
        elif not self.init2_run:
            self.init2()
        v = object.__new__(type(self))
        v.__dict__.update(self.__dict__)
        v._w_scale = float(self._w) if np.ndim(self._w) == 0 else 1.0   # what _context() derives from the weight
        v._parameters = None
        v._coefficients = None
        v._cache = {}
        v._fd_last = None
        v._launch_id = 0
        v._submit = submit
        v._parent = self
        v.stats = dict(launches=0, evaluated=0, hits=0, misses=0)
        return v

    # -- weight function ------------------------------------------------------------------------
    @property
    def w(self):
        return self._w

    @w.setter
    def w(self, value):
        # The upload is deferred to the next evaluation, so the reference idiom "assign zeros, then
        # fill in place" (weight.py:8-11) works; call touch_weight() after any LATER in-place edit.
        self._w = value
        self.touch_weight()

    def touch_weight(self):
        """Tell the object that `w` was modified in place: re-upload it and drop cached results."""
        self._w_dirty = True
        self._cache.clear()
        self._fd_last = None

    # -- device context ---------------------------------------------------------------------------
    def _wanted_devices(self, evals):
        """CUDA ordinals the context should span for a batch of `evals` evaluations."""
        devs = self.devices
        if devs is None:
            devs = os.environ.get('OPTSTENCIL_DEVICES') or None
            if devs is not None and devs != 'all':
                devs = [int(t) for t in devs.split(',') if t.strip() != '']
        if devs is None:
            first = self.device if self.device is not None else int(os.environ.get('OPTSTENCIL_DEVICE', 0))
            if evals >= self.auto_multi_evals and _capi.device_count() > 1:
                n = _capi.device_count()
                return [(first + i) % n for i in range(n)]
            return [first]
        if isinstance(devs, str):
            return list(range(max(1, _capi.device_count())))
        return [int(d) for d in devs]

    def _context(self, evals=0.0):
        """The device context (created lazily; grown to a multi-GPU group when a batch is large enough)."""
        if self._parent is not None:
            return self._parent._context(evals)
        if not self.init2_run:
            self.init2()
        if self._ctx is not None and (self.devices is not None or evals >= self.auto_multi_evals):
            want = self._wanted_devices(evals)
            grow = len(want) > len(self._ctx.devices)         # automatic mode never shrinks back: the extra contexts
            if grow or (self.devices is not None and list(want) != list(self._ctx.devices)):   # are paid for once
                self._ctx.close()
                self._ctx = None
        if self._ctx is None:
            cos, s2, kc = self._tables
            want = self._wanted_devices(evals)
            self._ctx = _capi.Context(self.dim, cos, s2, kc, Y=self.Y, Z=self.Z, device=want[0],
                                      devices=want if len(want) > 1 else None)
            self._w_dirty = True
        if self._w_dirty:
            w = self._w
            self._w_scale = 1.0
            if np.ndim(w) == 0:
                # a scalar weight factors out of the sum; the kernel then runs with unit weights
                self._ctx.set_weight_unit()
                self._w_scale = float(w)
            else:
                self._ctx.set_weight(np.asarray(w, dtype=float))
            self._w_dirty = False
        return self._ctx

    # -- parameters / coefficients (dispersion.py:78-109) ------------------------------------------
    @property
    def parameters(self):
        return self._parameters

    @parameters.setter
    def parameters(self, parameters):
        if not self.init2_run:
            self.init2()
        self._parameters = np.array(parameters)   # a copy, like the reference (dispersion.py:98-99)
        self._coefficients = None                 # the record array is built when somebody asks for it

    @property
    def coefficients(self):
        if self._coefficients is None and self._parameters is not None:
            self._coefficients = self.stencil.coefficients(self._parameters)
        return self._coefficients

    def _full(self):
        return np.array(self.coefficients.tolist()[0], dtype=float)

    # -- batched evaluation (new) -----------------------------------------------------------------
    # cache entry: [S, Amin, Amax, nan, full coefficient row, gate]; gate = (dt_multiplier, stencil_ok,
    # dt_ok, usable) is filled on first use, because its ValueErrors must surface only for points a
    # caller actually asked about (prefetched finite-difference neighbours may be garbage).
    def _grid_points(self):
        return float(self.N) ** self.dim

    def _launch(self, C, **kw):
        """One launch for the coefficient rows C -> (S, Amin, Amax, nan).  A view hands the rows to its server
        instead (they then travel as one SEGMENT of a shared launch: same bits as the direct call)."""
        if self._submit is not None:
            return self._submit(C)
        return self._context(len(C) * self._grid_points()).objective_batch(C, **kw)

    def _raw_batch(self, P):
        """(B, dof) parameters -> (S, Amin, Amax, nan, C) via ONE launch; fills the cache."""
        C = self.stencil.coefficients_batch(P)
        # the precise series, whatever the size: these are the values a search differentiates
        S, Amin, Amax, nan = self._launch(C) if self._submit is not None else self._launch(C, fast=False)
        self.stats['launches'] += 1
        self.stats['evaluated'] += len(P)
        self._launch_id += 1
        lid = self._launch_id
        cache = self._cache
        if len(cache) + len(P) > self.cache_size:
            for key in list(cache)[:len(cache) // 2 + len(P)]:   # drop the oldest half
                del cache[key]
        Sl, lo, hi, nl = S.tolist(), Amin.tolist(), Amax.tolist(), nan.tolist()
        for i in range(len(P)):
            cache[P[i].tobytes()] = [Sl[i], lo[i], hi[i], nl[i], C[i], None, lid]
        return S, Amin, Amax, nan, C

    def _entry(self, parameters):
        """Cache entry of one parameter vector (evaluating it if needed); None for NaN input."""
        p = parameters
        if type(p) is not np.ndarray or p.dtype != np.float64 or p.ndim != 1 or not p.flags.c_contiguous:
            p = np.ascontiguousarray(parameters, dtype=float).ravel()
        key = p.tobytes()
        e = self._cache.get(key)
        if e is not None:
            self.stats['hits'] += 1
            return e
        if np.isnan(p).any():
            return None
        self.stats['misses'] += 1
        if self.fd_prefetch:
            # exactly the points scipy's approx_derivative('2-point', abs_step=FD_STEP) will ask for:
            # x0 + h_vecs[i] with h_vecs = diag(h)  (scipy/optimize/_numdiff.py, _dense_difference)
            # (row 0 is x itself, row i + 1 a copy of x with x[i] + h: scipy's own copies, bit for bit)
            n = len(p)
            idx = self._fd_index.get(n)
            if idx is None:
                idx = self._fd_index[n] = (np.arange(1, n + 1), np.arange(n))
            P = np.repeat(p[None, :], n + 1, axis=0)
            P[idx] += FD_STEP
        else:
            P = p[None, :].copy()
        self._raw_batch(P)
        return self._cache[key]

    def _gate(self, e):
        """(stencil_ok, dt_ok, usable) of a cache entry -- dispersion.py:284-307 / 360-385, 129-154."""
        g = e[5]
        if g is not None and g[0] == self.dt_multiplier:
            return g[1], g[2], g[3]
        S, Amin, Amax, nan, c = e[0], e[1], e[2], e[3], e[4]
        if not Amin < 0:                     # also taken when Amin is NaN, as `not a < 0` is
            sok = self._corner_min(c)        # min(b): a shape-(1,) array in the reference
        else:
            sok = Amin                       # a = np.min(sqrtarg): a NumPy scalar in the reference
        if sok != sok:
            raise ValueError("stencil_ok got NaN")
        if sok < 0:
            dtok, usable = sok, False
        else:
            dt = float(c[0])
            root = self.dx * float(np.sqrt(Amax))
            dtok = (self.dt_multiplier / root if root != 0.0 else float('inf')) - dt
            if dtok != dtok:
                raise ValueError("dt_ok got NaN")
            usable = not (dtok < 0 or dt <= 0)
        e[5] = (self.dt_multiplier, sok, dtok, usable)
        return sok, dtok, usable

    def evaluate_batch(self, P):
        """Rows of free parameters -> dict of arrays norm / stencil_ok / dt_ok (reference semantics,
        including the 1e6 and -1 sentinels), one launch for all finite rows.  Gating is vectorised;
        batches of up to `cache_size`/4 rows also warm the per-point cache of the scalar API."""
        P = np.ascontiguousarray(P, dtype=float)
        if P.ndim == 1:
            P = P.reshape(1, -1)
        B = len(P)
        out = dict(norm=np.full(B, 1e6), stencil_ok=np.full(B, -1.0), dt_ok=np.full(B, -1.0))
        ok = ~np.any(np.isnan(P), axis=1)
        if not np.any(ok):
            return out
        Pok = np.ascontiguousarray(P[ok])
        # Small batches warm the per-point cache of the scalar API (precise series, like every search launch).  A batch
        # that is seconds of CPU work per row is a throughput launch whatever its row count: 512 rows on 512^3 are
        # what one of eight ranks gets of the 4096-row benchmark batch.
        if (len(Pok) * 4 <= self.cache_size and self._submit is None
                and len(Pok) * self._grid_points() < self.throughput_evals):
            S, Amin, Amax, nan, C = self._raw_batch(Pok)
        else:
            C = self.stencil.coefficients_batch(Pok)
            # large batches are what a start-grid screening looks like: let the library pick the kernel variant
            # with the all-NaN shortcut when most rows violate the CFL limit (same results either way)
            S, Amin, Amax, nan = self._launch(C, scan='auto') if self._submit is None else self._launch(C)
            self.stats['launches'] += 1
            self.stats['evaluated'] += len(Pok)
        with np.errstate(divide='ignore', invalid='ignore'):
            corner = self._corner_min_batch(C)
            sok = np.where(Amin < 0, Amin, corner)          # `a if a < 0 else min(b)`, NaN-transparent
            if np.any(np.isnan(sok)):
                raise ValueError("stencil_ok got NaN")
            dtok = self.dt_multiplier / (self.dx * np.sqrt(Amax)) - C[:, 0]
            dtok = np.where(sok < 0, sok, dtok)
            if np.any(np.isnan(dtok)):
                raise ValueError("dt_ok got NaN")
        usable = ~((sok < 0) | (dtok < 0) | (C[:, 0] <= 0))
        if np.any(usable & (nan > 0)):
            raise ValueError('Encountered NaNs in omega for params', Pok[np.flatnonzero(usable & (nan > 0))[0]])
        idx = np.flatnonzero(ok)
        out['stencil_ok'][idx] = sok
        out['dt_ok'][idx] = dtok
        out['norm'][idx] = np.where(usable, S * (self._w_scale * self._norm_scale()), 1e6)
        return out

    def norm_batch(self, P):
        return self.evaluate_batch(P)['norm']

    def _norm_scale(self):
        return (np.pi / (self.N - 1)) ** self.dim * self.Y ** 2 * self.Z ** 2   # dispersion.py:253

    # -- the reference's scalar API ----------------------------------------------------------------
    def stencil_ok(self, parameters):
        """min(sqrtarg) when that is negative (a NumPy scalar, as np.min gives), else the smallest corner value
        (a shape-(1,) array, as min() over the record fields gives) -- dispersion.py:284-307, 360-385."""
        self.parameters = parameters
        e = self._entry(parameters)
        if e is None:
            return -1
        sok = self._gate(e)[0]
        return np.float64(sok) if e[1] < 0 else np.array([sok])

    def dt_ok(self, parameters):
        """CFL margin; a negative stencil_ok is passed through unchanged (dispersion.py:129-154)."""
        self.parameters = parameters
        e = self._entry(parameters)
        if e is None:
            return -1.
        sok, dtok, _ = self._gate(e)
        if sok < 0:
            return np.float64(sok) if e[1] < 0 else np.array([sok])
        return np.array([dtok])

    def omega(self, parameters, out=None):
        """omega(k) on the dense grid, or None when the constraints are violated (dispersion.py:164-196).
        `out` (not in the reference): an existing float64 array of the grid's shape to fill and return."""
        self.parameters = parameters
        e = self._entry(parameters)
        if e is None:
            return None                      # stencil_ok() is -1 for NaN input (dispersion.py:174-178)
        if not self._gate(e)[2]:
            return None
        kw = {} if out is None else dict(out=out)
        om, _, _, nan = self._context().export(e[4], _capi.SB200_X_OMEGA, **kw)
        if nan:
            raise ValueError('Encountered NaNs in omega for params', parameters)
        return om

    def norm(self, parameters):
        """Weighted squared deviation from omega = k; 1e6 outside the constraints (dispersion.py:235-255)."""
        e = self._entry(parameters)
        if e is None:
            return 1e6
        self.parameters = parameters
        if not self._gate(e)[2]:
            return 1e6
        if e[3]:
            raise ValueError('Encountered NaNs in omega for params', parameters)
        return np.float64(e[0] * self._w_scale * self._norm_scale())   # numpy scalar, as f.sum() gives

    # -- SciPy's forward differences, reproduced bit for bit (new) ----------------------------------
    # SLSQP without `jac` calls approx_derivative(fun, x, method='2-point', abs_step=FD_STEP) once for
    # the objective and once per constraint at every iterate (scipy/optimize/_slsqp_py.py, cjac_factory;
    # _differentiable_functions.py).  Its arithmetic (_numdiff.py, _dense_difference) is
    #     x1 = copy(x); x1[i] = x[i] + h;   J[i] = (f(x1) - f(x)) / ((x[i] + h) - x[i])
    # The three *_fd methods below return exactly that from the result cache (one launch serves all of
    # them), so handing them to SciPy as `jac` leaves every iterate of the search unchanged and only
    # removes ~100 us of Python per approx_derivative call.
    def _norm_of(self, e, p):
        if e is None or not self._gate(e)[2]:
            return 1e6
        if e[3]:
            raise ValueError('Encountered NaNs in omega for params', p)
        return e[0] * self._w_scale * self._norm_scale()

    def _stencil_ok_of(self, e, p):
        return -1 if e is None else self._gate(e)[0]

    def _dt_ok_of(self, e, p):
        return -1. if e is None else self._gate(e)[1]

    def _fd_points(self, x):
        """(p, dx, entries) of x and its n forward-difference neighbours; memoised for the last x, because SLSQP
        asks for the Jacobians of the objective and of every constraint at the same iterate.  dx is None
        where SciPy would switch to a relative step (or x is not finite): the caller then lets SciPy do it."""
        p = np.array(x, dtype=float).ravel()
        key = p.tobytes()
        last = self._fd_last
        if last is not None and last[0] == key and last[4] is self._cache:
            return last[1], last[2], last[3]
        n = len(p)
        ph = p + FD_STEP
        dx = ph - p
        if not np.all(dx) or not np.all(np.isfinite(p)):
            self._fd_last = (key, p, None, None, self._cache)
            return p, None, None
        P = np.repeat(p[None, :], n + 1, axis=0)
        P[np.arange(1, n + 1), np.arange(n)] = ph            # row 0: x; row i + 1: copy(x); x1[i] = x[i] + h
        cache = self._cache
        entries = [cache.get(P[i].tobytes()) for i in range(n + 1)]
        lid = entries[0][6] if entries[0] is not None else None
        if lid is None or any(e is None or e[6] != lid for e in entries):
            # x and its n neighbours in ONE launch: a forward difference never mixes the sums of two launches
            # (they may differ in their last bits), and a point cached without its neighbours -- a start taken
            # from a screening batch, prefetch switched off -- costs one launch, not one per neighbour
            self.stats['misses'] += 1
            self._raw_batch(P)
            entries = [cache[P[i].tobytes()] for i in range(n + 1)]
        else:
            self.stats['hits'] += n + 1
        self._fd_last = (key, p, dx.tolist(), entries, self._cache)
        self._fd_last_point = P[n]
        return p, self._fd_last[2], entries

    def _fd_jac(self, value_of, scalar_fun, x):
        p, dx, entries = self._fd_points(x)
        if dx is None:
            # SciPy switches to a relative step there (or the point is garbage): let it do the work
            from scipy.optimize._numdiff import approx_derivative
            return approx_derivative(scalar_fun, p, method='2-point', abs_step=FD_STEP)
        n = len(p)
        f0 = value_of(entries[0], p)
        J = np.empty(n)
        for i in range(n):
            e = entries[i + 1]
            if e is None or e[3]:                            # the value may raise: give it the right point
                x1 = p.copy()
                x1[i] = p[i] + FD_STEP
                fi = value_of(e, x1)
            else:
                fi = value_of(e, None)
            J[i] = (fi - f0) / dx[i]
        self.parameters = self._fd_last_point                # the last point the reference would have seen
        return J

    def norm_fd(self, parameters):
        """Forward-difference gradient of norm() exactly as SciPy's SLSQP would compute it."""
        return self._fd_jac(self._norm_of, self.norm, parameters)

    def stencil_ok_fd(self, parameters):
        return self._fd_jac(self._stencil_ok_of, self.stencil_ok, parameters)

    def dt_ok_fd(self, parameters):
        return self._fd_jac(self._dt_ok_of, self.dt_ok, parameters)

    @property
    def sqrtarg(self):
        return self._context().export(self._full(), _capi.SB200_X_SQRTARG)[0]

    @property
    def sqrtres(self):
        res, _, _, nan = self._context().export(self._full(), _capi.SB200_X_SQRTRES)
        if nan:
            print('This Function should not have been called')
            raise ValueError('Got NaNs in sqrt for params {}, stencil_ok {}'.format(
                self._parameters, self.stencil_ok(self._parameters)))
        return res

    # -- consumers of the dense omega (dispersion.py:199-232, 322-324, 402-413) --------------------
    def _velocity_fields(self, parameters, want):
        """omega and what the reference derives from it with numpy.gradient (dispersion.py:208-220,
        408-410), computed on the device in NumPy's own arithmetic.  None when omega() would be None."""
        self.parameters = parameters
        e = self._entry(parameters)
        if e is None or not self._gate(e)[2]:
            return None
        f = self._context().group_velocity(e[4], self.dks, want)
        if f['nan']:
            raise ValueError('Encountered NaNs in omega for params', parameters)
        if 'vg' in f:
            f['vg'][(slice(None, 3),) * self.dim] = 1
        if 'vph' in f:
            f['vph'].ravel()[0] = 1.
        return f

    def omega_output(self, fname, parameters):
        """gnuplot-style table: one block per kx index, columns kx ky [kz] k omega vph vg vgx vgy [vgz]."""
        f = self._velocity_fields(parameters, ("omega", "vgc", "vg", "vph", "k"))
        if f is None:
            raise ValueError('omega is None for params {}: outside the constraints'.format(parameters))
        cols = np.broadcast_arrays(*self.kappamesh, f['k'], f['omega'], f['vph'], f['vg'], *f['vgc'])
        names = ' '.join('kx ky kz'.split()[:self.dim]) + ' k omega vph vg ' + ' '.join('vgx vgy vgz'.split()[:self.dim])
        with open(fname, 'wb') as out:
            out.write(('#' + names + '\n').encode('us-ascii'))
            for ix in range(cols[0].shape[0]):
                np.savetxt(out, np.stack([np.ravel(col[ix]) for col in cols], axis=1))
                out.write(b'\n')

    def _corner_min(self, c):
        raise NotImplementedError


class Dispersion2D(Dispersion):
    """Two-dimensional dispersion relation."""

    dim = 2

    def __init__(self, Y, N, stencil, **kwargs):
        self.Y = Y
        self.N = N
        self.stencil = stencil
        super().__init__(**kwargs)

    def _corner_min(self, c):
        """sqrtarg at the corners kappa in {0,pi}^2 other than the origin (dispersion.py:294-297)."""
        dt, ax, ay, bxy, byx, dlx, dly = c
        Y2 = self.Y ** 2
        return min((ay + 2 * byx - dly) / Y2,
                   ax - 2 * bxy - dlx + (ay - 2 * byx - dly) / Y2,
                   ax + 2 * bxy - dlx)

    def _corner_min_batch(self, C):
        dt, ax, ay, bxy, byx, dlx, dly = C.T
        Y2 = self.Y ** 2
        b = [(ay + 2 * byx - dly) / Y2, ax - 2 * bxy - dlx + (ay - 2 * byx - dly) / Y2, ax + 2 * bxy - dlx]
        return _python_min(b)

    def omega_spline(self, parameters, s=0):
        omega = self.omega(parameters)
        return spinterp.RectBivariateSpline(self.kx[:, 0], self.ky[0, :], omega, s=s)


class Dispersion3D(Dispersion):
    """Three-dimensional dispersion relation."""

    dim = 3

    def __init__(self, Y, Z, N, stencil, **kwargs):
        self.Y = Y
        self.Z = Z
        self.N = N
        self.stencil = stencil
        super().__init__(**kwargs)

    def _corner_min(self, c):
        """sqrtarg at the seven corners kappa in {0,pi}^3 other than the origin (dispersion.py:369-375)."""
        dt, ax, ay, az, bxy, bxz, byx, byz, bzx, bzy, dlx, dly, dlz = c
        Y2, Z2 = self.Y ** 2, self.Z ** 2
        return min((az + 2.0 * bzx + 2.0 * bzy - dlz) / Z2,
                   (ay + 2.0 * byx + 2.0 * byz - dly) / Y2,
                   (ay + 2.0 * byx - 2.0 * byz - dly) / Y2 + (az + 2.0 * bzx - 2.0 * bzy - dlz) / Z2,
                   ax + 2.0 * bxy + 2.0 * bxz - dlx,
                   ax + 2.0 * bxy - 2.0 * bxz - dlx + (az - 2.0 * bzx + 2.0 * bzy - dlz) / Z2,
                   ax - 2.0 * bxy + 2.0 * bxz - dlx + (ay - 2.0 * byx + 2.0 * byz - dly) / Y2,
                   ax - 2.0 * bxy - 2.0 * bxz - dlx + (ay - 2.0 * byx - 2.0 * byz - dly) / Y2
                   + (az - 2.0 * bzx - 2.0 * bzy - dlz) / Z2)

    def _corner_min_batch(self, C):
        dt, ax, ay, az, bxy, bxz, byx, byz, bzx, bzy, dlx, dly, dlz = C.T
        Y2, Z2 = self.Y ** 2, self.Z ** 2
        b = [(az + 2.0 * bzx + 2.0 * bzy - dlz) / Z2,
             (ay + 2.0 * byx + 2.0 * byz - dly) / Y2,
             (ay + 2.0 * byx - 2.0 * byz - dly) / Y2 + (az + 2.0 * bzx - 2.0 * bzy - dlz) / Z2,
             ax + 2.0 * bxy + 2.0 * bxz - dlx,
             ax + 2.0 * bxy - 2.0 * bxz - dlx + (az - 2.0 * bzx + 2.0 * bzy - dlz) / Z2,
             ax - 2.0 * bxy + 2.0 * bxz - dlx + (ay - 2.0 * byx + 2.0 * byz - dly) / Y2,
             ax - 2.0 * bxy - 2.0 * bxz - dlx + (ay - 2.0 * byx - 2.0 * byz - dly) / Y2
             + (az - 2.0 * bzx - 2.0 * bzy - dlz) / Z2]
        return _python_min(b)

    def omega_interp(self, parameters):
        f = self._velocity_fields(parameters, ("omega", "vg"))
        axes = (self.kx[:, 0, 0], self.ky[0, :, 0], self.kz[0, 0, :])
        return spinterp.RegularGridInterpolator(axes, f['omega']), spinterp.RegularGridInterpolator(axes, f['vg'])
