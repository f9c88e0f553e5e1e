"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/oracle_dispersion.c (built by `make -C oracle`)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle_dispersion.so")
_dp = ctypes.POINTER(ctypes.c_double)


class _Grid(ctypes.Structure):
    _fields_ = [("dim", ctypes.c_int), ("n", ctypes.c_int * 3), ("cosk", _dp * 3), ("s2", _dp * 3), ("kcomp", _dp * 3),
                ("Y", ctypes.c_double), ("Z", ctypes.c_double), ("Y2", ctypes.c_double), ("Z2", ctypes.c_double),
                ("wkind", ctypes.c_int), ("w", _dp)]


def build(force=False):
    # rebuild on request or when missing; a copied tree may carry arbitrary mtimes, so an existing
    # library is only considered stale when the build container asks (force=True in __graft_entry__.build)
    if force or not os.path.exists(LIB):
        subprocess.run(["make", "-C", HERE, "-B", "liboracle_dispersion.so"], check=True, capture_output=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
        _lib.oracle_objective.argtypes = [ctypes.POINTER(_Grid), _dp, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                          _dp, _dp, _dp, ctypes.POINTER(ctypes.c_int64)]
        _lib.oracle_export.argtypes = [ctypes.POINTER(_Grid), _dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _dp]
        assert _lib.oracle_abi_version() == 1
    return _lib


class COracle:
    """cos / s2 / kcomp: per-axis 1-D tables (as for the C-ABI); w: None, (ny, nz) plane or dense array."""

    def __init__(self, dim, cos, s2, kcomp, Y=1.0, Z=1.0, w=None):
        self.dim = dim
        self._keep = [[np.ascontiguousarray(np.ravel(t[a]), dtype=float) for a in range(dim)] for t in (cos, s2, kcomp)]
        g = _Grid()
        g.dim, g.Y, g.Z = dim, float(Y), float(Z)
        dx = 1.0
        g.Y2, g.Z2 = float((Y * dx) ** 2), float((Z * dx) ** 2)   # the reference's expression, with its types
        self.shape = tuple(t.size for t in self._keep[0])
        for a in range(3):
            g.n[a] = self.shape[a] if a < dim else 1
        for a in range(dim):
            g.cosk[a], g.s2[a], g.kcomp[a] = (self._keep[i][a].ctypes.data_as(_dp) for i in range(3))
        g.wkind = 0
        if w is not None and np.ndim(w) > 0:
            w = np.asarray(w, dtype=float)
            if dim == 3 and w.ndim == 3 and w.shape[0] == 1:
                self._w = np.ascontiguousarray(np.broadcast_to(w, (1,) + self.shape[1:])[0])
                g.wkind = 1
            else:
                self._w = np.ascontiguousarray(np.broadcast_to(w, self.shape))
                g.wkind = 2
            g.w = self._w.ctypes.data_as(_dp)
        self._g = g

    def objective(self, coeffs, x0=0, x1=None):
        c = np.ascontiguousarray(coeffs, dtype=float)
        if c.ndim == 1:
            c = c.reshape(1, -1)
        B = len(c)
        x1 = self.shape[0] if x1 is None else x1
        S, lo, hi = np.empty(B), np.empty(B), np.empty(B)
        nan = np.empty(B, dtype=np.int64)
        lib().oracle_objective(ctypes.byref(self._g), c.ctypes.data_as(_dp), B, x0, x1, S.ctypes.data_as(_dp),
                               lo.ctypes.data_as(_dp), hi.ctypes.data_as(_dp),
                               nan.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
        return S, lo, hi, nan

    def export(self, coeffs, what=0, x0=0, x1=None):
        c = np.ascontiguousarray(coeffs, dtype=float)
        x1 = self.shape[0] if x1 is None else x1
        out = np.empty((x1 - x0,) + self.shape[1:])
        lib().oracle_export(ctypes.byref(self._g), c.ctypes.data_as(_dp), what, x0, x1, out.ctypes.data_as(_dp))
        return out
