"""CPU stand-in for optimize_stencil_b200._capi.Context, backed by the oracle.

TEST INFRASTRUCTURE ONLY.  It implements the device contract of include/stencil_b200.h
(un-gated S, Amin = min(sqrtarg U {+0}), Amax = max(sqrtarg U {+0}), NaN count; dense exports) with
oracle/np_oracle.py so that the HOST logic of the drop-in package (gating, caching, finite-difference
prefetch, the Optimize scan, sharding) can be exercised by `-m "not gpu"` tests.  The product never
imports this; on a GPU box the same tests run against the real CUDA context.
"""
import numpy as np

import np_oracle as O


class OracleContext:
    instances = 0

    def __init__(self, dim, cos_kappa, s2, kcomp, Y=1.0, Z=1.0, device=0, devices=None):
        self.dim = dim
        self.devices = list(devices) if devices else [device]
        self.ncoef = 7 if dim == 2 else 13
        self.grid = O.Grid.from_tables(dim, [np.ravel(c) for c in cos_kappa], [np.ravel(c) for c in s2],
                                       [np.ravel(c) for c in kcomp], Y, Z)
        self.shape = self.grid.shape
        self.w = 1.0
        self.launch_count = 0
        self.last_batch_ms = 0.0
        OracleContext.instances += 1

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        pass

    def set_weight_unit(self):
        self.w = 1.0

    def set_weight(self, w):
        self.w = np.array(np.broadcast_to(np.asarray(w, dtype=float), self.shape))

    def set_tiling(self, *a):
        pass

    def _slab(self, x_begin, x_end):
        if x_begin == 0 and (x_end is None or x_end == self.shape[0]):
            return self.grid, self.w
        g = self.grid
        sl = slice(x_begin, x_end)
        sub = O.Grid.from_tables(self.dim, [g.cos[0].ravel()[sl]] + [c.ravel() for c in g.cos[1:]],
                                 [g.s2[0].ravel()[sl]] + [c.ravel() for c in g.s2[1:]],
                                 [g.kcomp[0].ravel()[sl]] + [c.ravel() for c in g.kcomp[1:]], g.Y, g.Z)
        return sub, (self.w if np.ndim(self.w) == 0 else self.w[sl])

    def objective_batch(self, coeffs, x_begin=0, x_end=None, x_stride=1, seg_size=0, scan=False, shard="auto",
                        fast=None):
        # segments, the screening variant and sharding change nothing in the results contract
        assert x_stride == 1, "the oracle-backed context evaluates contiguous slabs only"
        c = np.asarray(coeffs, dtype=float)
        if c.ndim == 1:
            c = c.reshape(1, -1)
        if c.shape[1] != self.ncoef:
            raise ValueError("coeffs must have shape (B, %d)" % self.ncoef)
        grid, w = self._slab(x_begin, x_end)
        B = len(c)
        S, Amin, Amax, nan = np.empty(B), np.empty(B), np.empty(B), np.empty(B, dtype=np.int64)
        for b in range(B):
            s, amin, amax, nn = O.device_contract(grid, c[b], w)
            bad = np.isnan(amin) or np.isnan(amax)
            S[b], nan[b] = s, nn
            Amin[b] = np.nan if bad else (amin if amin < 0 else 0.0)
            Amax[b] = np.nan if bad else max(amax, 0.0)
        self.launch_count += 1
        return S, Amin, Amax, nan

    def export(self, coeffs, what=0, x_begin=0, x_end=None, out=None):
        grid, _ = self._slab(x_begin, x_end)
        c = np.asarray(coeffs, dtype=float)
        with np.errstate(invalid="ignore", divide="ignore"):
            A = np.array(np.broadcast_to(O.sqrtarg(grid, c), grid.shape))
            res = A if what == 1 else (np.sqrt(A) if what == 2 else O.omega_from_sqrtarg(A, c[0]))
        if out is not None:
            out[...] = res
        else:
            out = res
        amin, amax = A.min(), A.max()
        self.launch_count += 1
        return out, (amin if amin < 0 else 0.0), max(amax, 0.0), int(np.count_nonzero(np.isnan(out)))


    def group_velocity(self, coeffs, spacings, want=("omega", "vgc", "vg", "vph", "k")):
        grid = self.grid
        c = np.asarray(coeffs, dtype=float)
        with np.errstate(invalid="ignore", divide="ignore"):
            A = np.array(np.broadcast_to(O.sqrtarg(grid, c), grid.shape))
            om = O.omega_from_sqrtarg(A, c[0])
            vgs = np.gradient(om, *spacings, edge_order=2)
            f = dict(omega=om, vgc=tuple(vgs), vg=np.sqrt(sum(v ** 2 for v in vgs)), vph=om / grid.k,
                     k=np.array(np.broadcast_to(grid.k, grid.shape)), nan=int(np.count_nonzero(np.isnan(om))))
        self.launch_count += 1
        return {k: v for k, v in f.items() if k in want or k == "nan"}


def install(monkeypatch):
    """Route the drop-in package to the oracle-backed context (CPU tests only)."""
    from optimize_stencil_b200 import _capi
    monkeypatch.setattr(_capi, "Context", OracleContext)
