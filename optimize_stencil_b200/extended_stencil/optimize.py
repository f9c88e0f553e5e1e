"""Multi-start SLSQP driver (drop-in for the reference's extended_stencil/optimize.py).

SciPy still drives every search (`scipy.optimize.minimize(method='SLSQP')` with the reference's
constraints and options, optimize.py:136,142).  What changes is the data-parallel part:

  reference: mp.Pool(nproc).imap(_optimize_single, starts) -- one whole SLSQP run per CPU worker,
             every objective / constraint call a fresh ~25-pass NumPy sweep (optimize.py:169-190);
  here:      1. the start grid is screened in ONE batched launch (stencil_ok and the CFL time step of every
                start; optimize.py:118-134), infeasible starts never reach SciPy;
             2. the surviving searches run CONCURRENTLY in this process, one thread each, in lock-step: when
                every live search has asked for its next iterate (the iterate and its finite-difference
                neighbours), ONE launch serves them all -- each search's points travel as one segment of that
                launch, and the library guarantees a segment's results are bit-identical to a launch of its
                own (include/stencil_b200.h, sb200_objective_batch_ex), so every search is bit-identical to
                its stand-alone run (`lockstep = False`) whatever shares the launch;
             3. for a large start grid SciPy's own per-iteration Python time (~0.1 ms per iterate, serial under
                the interpreter lock) is what remains, so the searches are additionally dealt to worker PROCESSES
                (`workers`).  The workers run SciPy only -- they never touch CUDA.  This process keeps the device
                contexts (all visible GPUs as one device group) and serves them: every round it merges the blocks
                of ALL workers into one segmented launch per dispersion object, sharded over the GPUs at segment
                boundaries, and hands the slices back.  One Python process owns the devices (SURVEY.md 8e "process
                model"); several processes sharing a GPU without MPS time-slice it and lose more than they gain
                (measured: profiles/r02_scan_timing_1gpu.txt).  Starts are independent and deterministic and a
                segment's bits do not depend on its company, so the result does not depend on the number of
                workers, threads or GPUs.
             SLSQP is handed the forward differences it would take itself as `jac` callables and the
             constraint list stacked into one vector-valued constraint (`exact_fd_jac`: same points, same
             arithmetic, same rows -- the search is bit-identical, the Python overhead per iterate drops 3x).
             `--singlecore` keeps everything in this process (no worker processes), like the reference.

The printed lines keep the reference's format (optimize.py:71, 144, 191, 200-202).
"""
import contextlib
import itertools
import os
import sys
import threading
import time

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


_T0 = time.perf_counter()


def _trace(what):
    """OPTSTENCIL_TRACE=1: stage times on stderr (seconds since this module was imported)."""
    if os.environ.get('OPTSTENCIL_TRACE'):
        sys.stderr.write('[optimize %8.3f s] %s\n' % (time.perf_counter() - _T0, what))


class _Request:
    __slots__ = ("parent", "C", "result", "error", "lock")

    def __init__(self, parent, C):
        self.parent, self.C, self.result, self.error = parent, C, None, None
        self.lock = threading.Lock()
        self.lock.acquire()                      # released when it is this request's turn to continue


class _LockStep:
    """Runs many independent searches as cooperatively scheduled threads and merges their device work into rounds.

    Exactly ONE search thread runs at any time (a baton is passed explicitly), so the threads never fight over
    the interpreter lock -- with dozens of runnable threads CPython's forced GIL hand-overs cost more than the
    searches themselves.  A search that needs the objective queues its coefficient rows (`submit`) and passes the
    baton to the next search whose results are ready; when none is -- every live search waits for results -- the
    thread holding the baton runs the round: the queued blocks are concatenated per (dispersion object, block
    length) and evaluated with ONE segmented launch each; every block is one segment, i.e. bit-identical to a
    launch of its own.  What the rounds remove is the per-iterate launch + sync latency (~45 us each) and the GPU
    idling between the small launches of a sequential scan; SciPy's own Python time is unchanged."""

    pin = os.environ.get('OPTSTENCIL_PIN', '1') != '0'

    def __init__(self, evaluate, pin_index=-1):
        self.evaluate = evaluate     # [(key, seg, C), ...] -> [(S, Amin, Amax, nan), ...]: the device work of one round
        self.pin_index = pin_index   # which of the allowed cores the search threads share (-1: the last one)
        self._waiting = []           # requests queued for the next round
        self._ready = []             # requests whose results are in (or start tokens), in wake-up order
        self._idle = threading.Lock()
        self.rounds = 0
        self.launches = 0
        self.rows = 0
        self.device_seconds = 0.0    # wall time spent waiting for the device work over all rounds

    def submitter(self, key):
        def submit(C):
            req = _Request(key, C)
            self._waiting.append(req)
            self._pass_baton()
            req.lock.acquire()                   # immediately, if this thread ran the round and is first in line
            if req.error is not None:
                raise req.error
            return req.result
        return submit

    def _pass_baton(self):
        """Called by the running thread just before it blocks or exits."""
        if not self._ready and self._waiting:
            self._round()
        if self._ready:
            self._ready.pop().lock.release()
        else:
            self._idle.release()                 # nothing ready, nothing waiting: every thread is done

    def _round(self):
        batch, self._waiting = self._waiting, []
        try:
            groups = {}
            for req in batch:
                groups.setdefault((req.parent, len(req.C)), []).append(req)
            blocks = [(key, seg, np.concatenate([r.C for r in reqs]) if len(reqs) > 1 else reqs[0].C)
                      for (key, seg), reqs in groups.items()]
            t0 = time.perf_counter()
            answers = self.evaluate(blocks)
            self.device_seconds += time.perf_counter() - t0
            for ((key, seg), reqs), (S, Amin, Amax, nan) in zip(groups.items(), answers):
                self.launches += 1
                self.rows += seg * len(reqs)
                for n, r in enumerate(reqs):
                    sl = slice(n * seg, (n + 1) * seg)
                    r.result = (S[sl], Amin[sl], Amax[sl], nan[sl])
            self.rounds += 1
        except BaseException as exc:             # a failed launch fails every search that waited for it
            for r in batch:
                if r.result is None:
                    r.error = exc
        batch.reverse()                          # popped from the end: first come, first served
        self._ready = batch

    def run(self, jobs, nthreads):
        """jobs: callables; returns their results in order (the first exception of any job is re-raised)."""
        results = [None] * len(jobs)
        errors = []
        nxt = [0]
        nthreads = max(1, min(nthreads, len(jobs)))
        tokens = [_Request(None, None) for _ in range(nthreads)]

        def worker(token):
            if core is not None:
                try:
                    os.sched_setaffinity(0, {core})      # 0 = the calling thread
                except OSError:
                    pass
            token.lock.acquire()                 # wait for the baton
            try:
                while not errors:
                    n = nxt[0]
                    nxt[0] += 1
                    if n >= len(jobs):
                        break
                    try:
                        results[n] = jobs[n]()
                    except BaseException as exc:
                        errors.append(exc)
            finally:
                self._pass_baton()

        # Only one search thread runs at a time, so they all share ONE core: a baton hand-over is then a same-core
        # context switch with warm caches (~5 us) instead of waking an idle core (100-300 us measured).
        core = None
        if self.pin and hasattr(os, 'sched_setaffinity'):
            allowed = sorted(os.sched_getaffinity(0))
            core = allowed[self.pin_index % len(allowed)] if allowed else None
        self._idle.acquire()
        threads = [threading.Thread(target=worker, args=(tok,), daemon=True) for tok in tokens]
        for t in threads:
            t.start()
        self._ready = tokens[::-1]
        self._ready.pop().lock.release()         # the first baton
        self._idle.acquire()                     # released by the last thread to finish
        self._idle.release()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return results


@contextlib.contextmanager
def _single_threaded_blas():
    """SLSQP's linear algebra is a handful of tiny LAPACK calls per iterate.  Their bits depend on the number of BLAS
    threads (measured: OMP_NUM_THREADS=1 against the default changes the printed optimum of a scan in the 10th
    digit, with the objective held bit-identical), and dozens of worker processes with one BLAS thread pool each
    oversubscribe the host.  One thread during a scan: the same trajectories on every machine and in every mode."""
    if os.environ.get('OPTSTENCIL_BLAS_THREADS', '1') != '1':      # anything else: leave the BLAS thread pools alone
        yield
        return
    try:
        from threadpoolctl import threadpool_limits
    except ImportError:
        yield
        return
    with threadpool_limits(limits=1, user_api='blas'):
        yield


def _launch_blocks(parents, blocks):
    """The device work of one round: one segmented launch per (dispersion object, block length)."""
    out = []
    for key, seg, C in blocks:
        parent = parents[key]
        ctx = parent._context(len(C) * parent._grid_points())
        out.append(ctx.objective_batch(C, seg_size=seg if len(C) > seg else 0, fast=False))
    return out


def _scipy_worker(conn, opt, rank):
    """Entry point of a worker process: its share of the searches, lock-stepped among themselves; every round's
    blocks go to the process that owns the devices (`Optimize._serve`) and come back evaluated.  No CUDA here.
    The share arrives over the pipe: the workers are started (and import NumPy / SciPy) while the owner still
    creates its device contexts and screens the start grid."""
    try:
        starts, screen = conn.recv()
        def remote(blocks):
            conn.send(('eval', blocks))
            answer = conn.recv()
            if isinstance(answer, BaseException):
                raise answer
            return answer
        with _single_threaded_blas():
            out = opt._scan_local(starts, screen, pin_index=rank, evaluate=remote)
        conn.send(('done', out, opt.scan_stats))
    except BaseException as exc:
        try:
            conn.send(('error', exc))
        except Exception:
            conn.send(('error', RuntimeError(repr(exc))))
    finally:
        conn.close()


class Optimize:
    """Holds the low- and high-resolution dispersion objects and runs the start-grid scan."""

    # Hand SciPy the finite differences it would take anyway (same points, same arithmetic, from the
    # result cache).  The search is bit-identical either way; OPTSTENCIL_FD_JAC=0 lets SciPy do them.
    exact_fd_jac = os.environ.get('OPTSTENCIL_FD_JAC', '1') != '0'
    # Run the searches of a scan concurrently, one launch per round (bit-identical to running them one by one).
    lockstep = os.environ.get('OPTSTENCIL_LOCKSTEP', '1') != '0'
    lockstep_threads = int(os.environ.get('OPTSTENCIL_THREADS', 256))
    # Worker processes for large start grids: None = automatic (none below `workers_min_starts` searches, else one
    # per `starts_per_worker` searches up to the physical cores), 0 = never, N = exactly N.
    workers = None
    serve_all_at_once = os.environ.get('OPTSTENCIL_SERVE', 'ready') == 'all'
    workers_min_starts = 512
    starts_per_worker = 128

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
        self.scan_stats = {}

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
    def _optimize_single(self, x0, low=None, high=None, screened=None):
        """One search.  `low` / `high`: the dispersion objects (or per-search views of them) to run it on;
        `screened`: (dt_ok at dt = 0, stencil_ok) of this start from the batched screening, else they are
        evaluated here like the reference does."""
        low = self.dispersion if low is None else low
        high = self.dispersion_high if high is None else high
        x0 = list(x0)
        if x0[0] is None:
            x0[0] = 0
            dt_ok = np.asarray(low.dt_ok(x0)).item() if screened is None else float(screened[0])
            if dt_ok < 0:
                return x0, None, float('inf')
            x0[0] = max(min(dt_ok, self.dtmax), self.dtmin)
        x0 = np.asarray(x0, dtype=float)
        # sqrtarg does not depend on dt, so the screening's stencil_ok (taken at dt = 0 when dt follows the CFL
        # rule) is the value the reference computes here
        if (low.stencil_ok(x0) if screened is None else screened[1]) < 0:
            return x0, None, float('inf')
        cons = self.constraints if low is self.dispersion else self._constraint_list(low)
        res = self._minimize(low, cons, x0)
        return x0, res, high.norm(res.x)

    def _optimize_final(self, x0):
        with _single_threaded_blas():
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
        """stencil_ok and dt_ok of every start from ONE batched launch (replaces the per-start NumPy sweeps the
        pool workers did first: optimize.py:118-134).  Rows whose dt follows the CFL rule are evaluated at
        dt = 0, as the reference does; both constraints derive from the bit-exact grid extrema, so they are
        the values the scalar API returns.  -> array [n_starts, 2] = (dt_ok, stencil_ok)."""
        if not starts:
            return np.empty((0, 2))
        P = np.asarray([[0.0 if v is None else v for v in s] for s in starts], dtype=float)
        out = self.dispersion.evaluate_batch(P)
        return np.stack([out['dt_ok'], out['stencil_ok']], axis=1)

    def _scan_local(self, starts, screen, pin_index=-1, evaluate=None):
        """Run the searches of `starts` in this process -> list of (x0, res, norm) in order.  `evaluate`: where a
        round's blocks are evaluated (None: on this process's device contexts)."""
        low, high = self.dispersion, self.dispersion_high
        lockstep = (self.lockstep and len(starts) >= 2) or evaluate is not None
        parents = dict(low=low, high=high)
        if evaluate is None:
            low._context(), high._context()                    # contexts and weights before any thread starts
            evaluate = lambda blocks: _launch_blocks(parents, blocks)   # noqa: E731
        server = _LockStep(evaluate, pin_index) if lockstep else None
        sub_low = server.submitter('low') if lockstep else None
        sub_high = sub_low if high is low else (server.submitter('high') if lockstep else None)

        def job(s, sc):
            def run():
                vl = low.view(sub_low)
                if high is low:
                    vh = vl
                else:
                    vh = high.view(sub_high)
                    vh.fd_prefetch = False                     # the high-resolution object only scores the result
                return self._optimize_single(s, vl, vh, sc), vl.stats, (None if vh is vl else vh.stats)
            return run

        jobs = [job(s, sc) for s, sc in zip(starts, screen)]
        done = server.run(jobs, self.lockstep_threads) if lockstep else [j() for j in jobs]
        for _, sl, sh in done:                                 # the views' counters end up on the shared objects
            for d, st in ((low, sl), (high, sh)):
                if st is not None:
                    for key, v in st.items():
                        d.stats[key] += v
        if lockstep:
            self.scan_stats = dict(rounds=server.rounds, launches=server.launches, rows=server.rows,
                                   device_seconds=round(server.device_seconds, 4))
        return [r for r, _, _ in done]

    def _worker_count(self, nstarts):
        w = getattr(self.args, 'workers', None)
        if w is None:
            w = os.environ.get('OPTSTENCIL_WORKERS')
        if w is None:
            w = self.workers
        if getattr(self.args, 'singlecore', False) and w is None:
            return 0
        if w is None:
            if nstarts < self.workers_min_starts:
                return 0
            try:
                import psutil
                cores = psutil.cpu_count(logical=False) or os.cpu_count() or 1
            except ImportError:
                cores = os.cpu_count() or 1
            w = min(cores - 1, max(1, nstarts // self.starts_per_worker))    # one core serves the devices
        w = int(w)
        return w if w > 1 and nstarts >= 2 * w else 0

    def _serve(self, conns):
        """Serve the worker processes until all are done: every round, the blocks of the workers that are waiting are
        merged per (dispersion object, block length) into one segmented launch (each block keeps its own segments,
        so nothing depends on the merging or on who shares a launch), and every worker gets its slices back."""
        parents = dict(low=self.dispersion, high=self.dispersion_high)
        active = dict(enumerate(conns))
        results, stats = {}, dict(rounds=0, launches=0, rows=0, device_seconds=0.0)
        failure = None
        round_log = [] if os.environ.get('OPTSTENCIL_ROUND_LOG') else None      # (round, rows, workers, device s, t)
        t_serve = time.perf_counter()
        from multiprocessing.connection import wait
        owner = {id(c): w for w, c in active.items()}
        while active:
            pending = {}
            # Whoever has handed in its blocks is served at once: while the devices work on them, the other workers
            # run their SLSQP steps, and what arrives meanwhile forms the next launch (a busy device batches by
            # itself, an idle one answers immediately).  `serve_all_at_once` waits for every live worker instead.
            ready = active if self.serve_all_at_once else {owner[id(c)]: c for c in wait(list(active.values()))}
            for w, c in sorted(ready.items()):
                try:
                    msg = c.recv()
                except EOFError:
                    msg = ('error', RuntimeError('worker %d died' % w))
                if msg[0] == 'eval':
                    pending[w] = msg[1]
                else:
                    del active[w]
                    if msg[0] == 'done':
                        results[w] = msg[1]
                    elif failure is None:
                        failure = msg[1]
            if not pending:
                continue
            merged = {}
            for w, blocks in pending.items():
                for n, (key, seg, C) in enumerate(blocks):
                    merged.setdefault((key, seg), []).append((w, n, C))
            replies = {w: [None] * len(blocks) for w, blocks in pending.items()}
            try:
                if failure is not None:
                    raise failure
                t0 = time.perf_counter()
                for (key, seg), items in merged.items():
                    C = np.concatenate([c for _, _, c in items]) if len(items) > 1 else items[0][2]
                    (S, Amin, Amax, nan), = _launch_blocks(parents, [(key, seg, C)])
                    stats['launches'] += 1
                    stats['rows'] += len(C)
                    at = 0
                    for w, n, c in items:
                        sl = slice(at, at + len(c))
                        replies[w][n] = (S[sl], Amin[sl], Amax[sl], nan[sl])
                        at += len(c)
                stats['rounds'] += 1
                stats['device_seconds'] += time.perf_counter() - t0
                if round_log is not None:
                    round_log.append((stats['rounds'], sum(len(c) for it in merged.values() for _, _, c in it),
                                      len(pending), time.perf_counter() - t0, time.perf_counter() - t_serve))
            except BaseException as exc:                       # fail every waiting worker; they report and exit
                if failure is None:
                    failure = exc
                replies = {w: exc for w in pending}
            for w in pending:
                active[w].send(replies[w])
        if failure is not None:
            raise failure
        stats['device_seconds'] = round(stats['device_seconds'], 4)
        if round_log is not None:
            with open(os.environ['OPTSTENCIL_ROUND_LOG'], 'w') as f:
                f.write('# round rows live_workers device_seconds seconds_since_start\n')
                for rec in round_log:
                    f.write('%d %d %d %.6f %.4f\n' % rec)
        return results, stats

    def _start_workers(self, nstarts):
        """Spawn the SciPy worker processes for a scan of `nstarts` searches (None: the scan stays in this process)."""
        workers = self._worker_count(nstarts)
        if workers <= 1:
            return None
        # Spawned by default: a fresh interpreter per worker costs 0.2 s of the scan's wall time (they start while this
        # process creates its CUDA context) and is safe whatever threads the caller already runs.  Forking
        # (OPTSTENCIL_START_METHOD=fork) is honoured only while this process has not called into CUDA; the workers
        # never touch CUDA themselves.
        from .. import _capi
        method = os.environ.get('OPTSTENCIL_START_METHOD', 'spawn')
        if method == 'fork' and (not hasattr(os, 'fork') or _capi.cuda_touched()):
            print('optimize: not forking a process that has called into CUDA; spawning the workers instead')
            method = 'spawn'
        if method == 'spawn' and not hasattr(sys.modules.get('__main__'), '__file__'):
            print('optimize: worker processes need an importable __main__ (not stdin / REPL); running in-process')
            return None
        import multiprocessing as mp
        mpc = mp.get_context(method)
        procs, conns = [], []
        for w in range(workers):
            parent_end, child_end = mpc.Pipe()
            p = mpc.Process(target=_scipy_worker, args=(child_end, self, w), daemon=True)
            p.start()
            child_end.close()
            procs.append(p); conns.append(parent_end)
        self._start_method = method
        return procs, conns, time.perf_counter()

    def _scan(self, starts, screen, pool=None):
        if pool is None:
            with _single_threaded_blas():
                return self._scan_local(starts, screen)
        from .. import _capi
        procs, conns, t_start = pool
        workers = len(procs)
        try:
            # The devices stay with this process; with more than one visible, the merged launches are sharded over all
            # of them (segments are bit-identical whatever the number of devices).
            if self.dispersion.devices is None and os.environ.get('OPTSTENCIL_DEVICES') is None and _capi.device_count() > 1:
                for d in (self.dispersion, self.dispersion_high):
                    d.devices = 'all'
            self.dispersion._context(), self.dispersion_high._context()
            _trace('device contexts ready')
            # The cost of a search depends on where its start lies (whole rows of the start grid are rejected at once,
            # others run a hundred iterates): deal the grid in a fixed pseudo-random order.  Every search is
            # independent and deterministic, so the result does not depend on the dealing.
            order = np.random.default_rng(len(starts)).permutation(len(starts))
            shares = [np.sort(order[w::workers]).tolist() for w in range(workers)]
            for c, idx in zip(conns, shares):
                c.send(([starts[i] for i in idx], screen[idx]))
            _trace('shares dealt')
            by_worker, stats = self._serve(conns)
            _trace('workers served')
        finally:
            for c in conns:
                c.close()
            for p in procs:
                p.join(timeout=30)
                if p.is_alive():
                    p.terminate()
            _trace('workers joined')
        results = [None] * len(starts)
        for w, idx in enumerate(shares):
            for i, r in zip(idx, by_worker[w]):
                results[i] = r
        self.scan_stats = dict(stats, workers=workers, start_method=self._start_method,
                               devices=len(self.dispersion._context().devices),
                               wall_seconds=round(time.perf_counter() - t_start, 3))
        return results

    def optimize(self):
        """Scan the grid of starting values, keep the best local optimum, polish it at high resolution."""
        starts = self.start_grid()
        _trace('start grid of %d points' % len(starts))
        pool = self._start_workers(len(starts))     # first: they start up while this process screens the grid
        _trace('workers started' if pool else 'no workers')
        try:
            screen = self._prescreen(starts)
            _trace('start grid screened')
        except BaseException:
            if pool is not None:
                for p in pool[0]:
                    p.terminate()
            raise
        best = None
        for x0, res, norm in self._scan(starts, screen, pool):
            print('x0={}, x={}, norm={}'.format(x0, res.x if res else None, norm))
            if best is None or norm < best[2]:
                best = (x0, res, norm)
        sys.stdout.flush()
        _trace('scan done: %s' % (self.scan_stats,))
        bestx0, bestres, bestnorm = best
        if self.args.Ngrid_high != self.args.Ngrid_low:
            print()
            print('using this result as starting point for high-res optimization:')
            print('x0={},x={}, norm={}'.format(bestx0, bestres.x, bestnorm))
            bestres, bestnorm = self._optimize_final(bestres.x)
        _trace('final optimisation done')
        return bestres.x, bestnorm

    def omega_output(self, fname, x):
        return self.dispersion_high.omega_output(fname, x)
