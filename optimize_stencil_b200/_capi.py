"""ctypes binding of include/stencil_b200.h (libstencil_b200.so).

This is the ONLY route from Python to the objective: there is no NumPy or CPU fallback.  If the
shared library is missing, was not built for this machine, or no CUDA device is usable, the calls
below raise `StencilB200Error` -- loudly, by design.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SB200_LIB") or os.path.join(_HERE, "csrc", "libstencil_b200.so")  # env: experiments

SB200_W_UNIT, SB200_W_PLANE, SB200_W_DENSE = 0, 1, 2
SB200_X_OMEGA, SB200_X_SQRTARG, SB200_X_SQRTRES = 0, 1, 2

EXPORTED_SYMBOLS = (
    "sb200_abi_version", "sb200_device_count", "sb200_create", "sb200_destroy", "sb200_last_error",
    "sb200_set_weight", "sb200_objective_batch", "sb200_objective_batch_device", "sb200_export",
    "sb200_export_device", "sb200_group_velocity", "sb200_get_stream", "sb200_launch_count", "sb200_last_batch_ms", "sb200_set_tiling", "sb200_fp64_peak",
)


class StencilB200Error(RuntimeError):
    pass


class _GridDesc(ctypes.Structure):
    _fields_ = [("dim", ctypes.c_int32), ("n", ctypes.c_int32 * 3),
                ("cos_kappa", ctypes.POINTER(ctypes.c_double) * 3),
                ("s2", ctypes.POINTER(ctypes.c_double) * 3),
                ("kcomp", ctypes.POINTER(ctypes.c_double) * 3),
                ("Y", ctypes.c_double), ("Z", ctypes.c_double), ("Y2", ctypes.c_double), ("Z2", ctypes.c_double),
                ("device", ctypes.c_int32)]


_lib = None


def load_library():
    """dlopen the in-tree shared library and declare the prototypes.  Raises when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise StencilB200Error(
            "libstencil_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `python -m optimize_stencil_b200._build`. There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    dp = ctypes.POINTER(ctypes.c_double)
    i64p = ctypes.POINTER(ctypes.c_int64)
    vp = ctypes.c_void_p
    lib.sb200_abi_version.restype = ctypes.c_int
    lib.sb200_device_count.argtypes = [ctypes.POINTER(ctypes.c_int32)]
    lib.sb200_create.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(_GridDesc)]
    lib.sb200_destroy.argtypes = [vp]
    lib.sb200_destroy.restype = None
    lib.sb200_last_error.argtypes = [vp]
    lib.sb200_last_error.restype = ctypes.c_char_p
    lib.sb200_set_weight.argtypes = [vp, ctypes.c_int32, dp, ctypes.c_int64]
    # the hot call takes plain addresses: ndarray.ctypes.data_as() costs 3 us per pointer, five of them per call
    lib.sb200_objective_batch.argtypes = [vp, vp, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, vp, vp, vp, vp]
    lib.sb200_objective_batch_device.argtypes = [vp, vp, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, vp, vp]
    lib.sb200_export.argtypes = [vp, dp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, dp, dp, dp, i64p]
    lib.sb200_export_device.argtypes = [vp, dp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, vp, vp, vp]
    lib.sb200_group_velocity.argtypes = [vp, dp, dp, dp, dp, dp, dp, dp, i64p]
    lib.sb200_get_stream.argtypes = [vp, ctypes.POINTER(vp)]
    lib.sb200_launch_count.argtypes = [vp, i64p]
    lib.sb200_last_batch_ms.argtypes = [vp, dp]
    lib.sb200_set_tiling.argtypes = [vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]
    lib.sb200_fp64_peak.argtypes = [ctypes.c_int32, ctypes.c_double, dp]
    for name in EXPORTED_SYMBOLS:
        getattr(lib, name)  # AttributeError here == header and library disagree
        if name not in ("sb200_destroy", "sb200_last_error"):
            getattr(lib, name).restype = ctypes.c_int
    if lib.sb200_abi_version() != 1:
        raise StencilB200Error("libstencil_b200.so has ABI version %d, expected 1" % lib.sb200_abi_version())
    _lib = lib
    return lib


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != shape:
        raise ValueError("expected shape %s, got %s" % (shape, a.shape))
    return a


def device_count():
    lib = load_library()
    n = ctypes.c_int32(0)
    if lib.sb200_device_count(ctypes.byref(n)) != 0:
        return 0
    return n.value


def fp64_peak_tflops(device=0, seconds=0.25):
    lib = load_library()
    out = ctypes.c_double(0.0)
    rc = lib.sb200_fp64_peak(device, seconds, ctypes.byref(out))
    if rc != 0:
        raise StencilB200Error("sb200_fp64_peak failed (%d): %s" % (rc, lib.sb200_last_error(None).decode()))
    return out.value


class Context:
    """One device-resident k-grid (sb200_ctx).  Not thread-safe; one per host thread.

    cos_kappa, s2, kcomp: per-axis 1-D tables exactly as the reference's init2 computes them
    (dispersion.py:268-281, 339-357) -- `dim` arrays each.
    """

    def __init__(self, dim, cos_kappa, s2, kcomp, Y=1.0, Z=1.0, device=0):
        self._lib = load_library()
        self._h = ctypes.c_void_p(None)
        self.dim = int(dim)
        if self.dim not in (2, 3) or not (len(cos_kappa) == len(s2) == len(kcomp) == self.dim):
            raise ValueError("dim must be 2 or 3 with one table triple per axis")
        self.ncoef = 7 if self.dim == 2 else 13
        keep = []
        d = _GridDesc()
        d.dim = self.dim
        shape = []
        for a in range(self.dim):
            tabs = [_f64(np.ravel(t)) for t in (cos_kappa[a], s2[a], kcomp[a])]
            if not (tabs[0].shape == tabs[1].shape == tabs[2].shape):
                raise ValueError("axis %d tables have different lengths" % a)
            keep.append(tabs)
            d.n[a] = tabs[0].shape[0]
            shape.append(tabs[0].shape[0])
            d.cos_kappa[a], d.s2[a], d.kcomp[a] = (_dptr(t) for t in tabs)
        d.Y, d.Z, d.device = float(Y), float(Z), int(device)
        # the divisors of dispersion.py:319, 398 with the caller's own types: `**2` on a Python float or a
        # NumPy scalar is libm/NumPy pow, which is not always the correctly rounded Y*Y
        dx = 1.0
        d.Y2, d.Z2 = float((Y * dx) ** 2), float((Z * dx) ** 2)
        self.shape = tuple(shape)
        rc = self._lib.sb200_create(ctypes.byref(self._h), ctypes.byref(d))
        if rc != 0:
            self._h = ctypes.c_void_p(None)
            raise StencilB200Error("sb200_create failed (%d): %s" % (rc, self._lib.sb200_last_error(None).decode()))

    # -- plumbing -------------------------------------------------------------------------------
    def _check(self, rc, what):
        if rc != 0:
            raise StencilB200Error("%s failed (%d): %s" % (what, rc, self._lib.sb200_last_error(self._h).decode()))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.sb200_destroy(self._h)
            self._h = ctypes.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- weights --------------------------------------------------------------------------------
    def set_weight_unit(self):
        self._check(self._lib.sb200_set_weight(self._h, SB200_W_UNIT, None, 0), "sb200_set_weight")

    def set_weight(self, w):
        """w: scalar-free array broadcastable to the grid.  Arrays that do not depend on the first
        axis of a 3-D grid go to the device as an (ny, nz) plane; everything else dense."""
        w = np.asarray(w, dtype=np.float64)
        x_independent = w.ndim == 3 and (w.shape[0] == 1 or (w.shape[0] == self.shape[0] and w.strides[0] == 0))
        if self.dim == 3 and x_independent:
            plane = _f64(np.broadcast_to(w[:1], (1,) + self.shape[1:])[0])
            self._check(self._lib.sb200_set_weight(self._h, SB200_W_PLANE, _dptr(plane), plane.size), "sb200_set_weight")
        else:
            dense = _f64(np.broadcast_to(w, self.shape))
            self._check(self._lib.sb200_set_weight(self._h, SB200_W_DENSE, _dptr(dense), dense.size), "sb200_set_weight")

    # -- the objective --------------------------------------------------------------------------
    def objective_batch(self, coeffs, x_begin=0, x_end=None):
        """coeffs: (B, ncoef) full coefficient vectors -> (S, Amin, Amax, nan) arrays of length B."""
        c = _f64(coeffs)
        if c.ndim == 1:
            c = c.reshape(1, -1)
        if c.ndim != 2 or c.shape[1] != self.ncoef:
            raise ValueError("coeffs must have shape (B, %d), got %s" % (self.ncoef, c.shape))
        B = c.shape[0]
        x_end = self.shape[0] if x_end is None else x_end
        out = np.empty((4, B))                       # S | Amin | Amax | nan (int64 in the same 8-byte slots)
        base, row = out.ctypes.data, 8 * B
        rc = self._lib.sb200_objective_batch(self._h, c.ctypes.data, B, int(x_begin), int(x_end), base, base + row,
                                             base + 2 * row, base + 3 * row)
        self._check(rc, "sb200_objective_batch")
        return out[0], out[1], out[2], out[3].view(np.int64)

    def objective_batch_device(self, d_coeffs_ptr, nbatch, d_out_ptr, x_begin=0, x_end=None, stream=None):
        """Device-pointer variant (torch tensors' .data_ptr()); asynchronous on `stream`, a raw
        cudaStream_t value (0 / None = CUDA's default stream, i.e. torch's default stream)."""
        x_end = self.shape[0] if x_end is None else x_end
        rc = self._lib.sb200_objective_batch_device(self._h, ctypes.c_void_p(d_coeffs_ptr), int(nbatch), int(x_begin),
                                                    int(x_end), ctypes.c_void_p(d_out_ptr),
                                                    ctypes.c_void_p(stream or 0))
        self._check(rc, "sb200_objective_batch_device")

    def export(self, coeffs, what=SB200_X_OMEGA, x_begin=0, x_end=None):
        """Dense omega / sqrtarg / sqrtres for one coefficient vector -> (array, Amin, Amax, nan)."""
        c = _f64(coeffs, (self.ncoef,))
        x_end = self.shape[0] if x_end is None else x_end
        out = np.empty((x_end - x_begin,) + self.shape[1:], dtype=np.float64)
        amin, amax, nan = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
        rc = self._lib.sb200_export(self._h, _dptr(c), int(what), int(x_begin), int(x_end), _dptr(out),
                                    ctypes.byref(amin), ctypes.byref(amax), ctypes.byref(nan))
        self._check(rc, "sb200_export")
        return out, amin.value, amax.value, nan.value

    def group_velocity(self, coeffs, spacings, want=("omega", "vgc", "vg", "vph", "k")):
        """omega, numpy.gradient(omega, *spacings, edge_order=2), |vg|, vph = omega/k and k on the device;
        returns a dict of host arrays (vgc: tuple of `dim` arrays) plus 'nan' (NaN count of omega)."""
        c = _f64(coeffs, (self.ncoef,))
        h = _f64(np.asarray(spacings, dtype=float), (self.dim,))
        bufs = {name: np.empty(((self.dim,) if name == "vgc" else ()) + self.shape) for name in want}
        ptr = lambda name: _dptr(bufs[name]) if name in bufs else None
        nan = ctypes.c_int64()
        rc = self._lib.sb200_group_velocity(self._h, _dptr(c), _dptr(h), ptr("omega"), ptr("vgc"), ptr("vg"), ptr("vph"),
                                            ptr("k"), ctypes.byref(nan))
        self._check(rc, "sb200_group_velocity")
        if "vgc" in bufs:
            bufs["vgc"] = tuple(bufs["vgc"][a] for a in range(self.dim))
        bufs["nan"] = nan.value
        return bufs

    def export_device(self, coeffs, d_out_ptr, what=SB200_X_OMEGA, x_begin=0, x_end=None, d_stats_ptr=None, stream=None):
        c = _f64(coeffs, (self.ncoef,))
        x_end = self.shape[0] if x_end is None else x_end
        rc = self._lib.sb200_export_device(self._h, _dptr(c), int(what), int(x_begin), int(x_end),
                                           ctypes.c_void_p(d_out_ptr),
                                           ctypes.c_void_p(d_stats_ptr) if d_stats_ptr else None,
                                           ctypes.c_void_p(stream or 0))
        self._check(rc, "sb200_export_device")

    # -- bookkeeping ----------------------------------------------------------------------------
    def set_tiling(self, cand_per_thread=0, tile_y=0, tile_x=0):
        self._check(self._lib.sb200_set_tiling(self._h, cand_per_thread, tile_y, tile_x), "sb200_set_tiling")

    @property
    def stream(self):
        """The context's own cudaStream_t as an integer."""
        st = ctypes.c_void_p(None)
        self._check(self._lib.sb200_get_stream(self._h, ctypes.byref(st)), "sb200_get_stream")
        return st.value or 0

    @property
    def launch_count(self):
        n = ctypes.c_int64(0)
        self._check(self._lib.sb200_launch_count(self._h, ctypes.byref(n)), "sb200_launch_count")
        return n.value

    @property
    def last_batch_ms(self):
        ms = ctypes.c_double(0.0)
        self._check(self._lib.sb200_last_batch_ms(self._h, ctypes.byref(ms)), "sb200_last_batch_ms")
        return ms.value
