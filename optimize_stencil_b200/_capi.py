"""ctypes binding of include/stencil_b200.h (libstencil_b200.so).

This is the ONLY route from Python to the objective: there is no NumPy or CPU fallback.  If the
shared library is missing, was not built for this machine, or no CUDA device is usable, the calls
below raise `StencilB200Error` -- loudly, by design.
"""
import ctypes
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SB200_LIB") or os.path.join(_HERE, "csrc", "libstencil_b200.so")  # env: experiments

SB200_W_UNIT, SB200_W_PLANE, SB200_W_DENSE = 0, 1, 2
SB200_X_OMEGA, SB200_X_SQRTARG, SB200_X_SQRTRES = 0, 1, 2

SB200_F_SCAN, SB200_F_AUTO_SCAN, SB200_F_FAST = 1, 2, 4
FAST_SERIES_EVALS = float(2 ** 28)   # automatic choice of the series: k-point x candidate evaluations per SEGMENT from which
                                     # a launch is a throughput launch (shorter series), not a search launch
SHARD_AUTO, SHARD_CANDIDATES, SHARD_SLABS, SHARD_NONE, SHARD_SLABS_CONTIGUOUS = 0, 1, 2, 3, 4
SHARD_MODES = {"auto": SHARD_AUTO, "candidates": SHARD_CANDIDATES, "slabs": SHARD_SLABS, "none": SHARD_NONE,
               "slabs-contiguous": SHARD_SLABS_CONTIGUOUS}
ABI_VERSION = 2

EXPORTED_SYMBOLS = (
    "sb200_abi_version", "sb200_device_count", "sb200_create", "sb200_destroy", "sb200_last_error",
    "sb200_set_weight", "sb200_objective_batch", "sb200_objective_batch_ex", "sb200_objective_batch_device",
    "sb200_objective_batch_device_ex", "sb200_combine_slabs_device", "sb200_export", "sb200_export_device",
    "sb200_group_velocity", "sb200_get_stream", "sb200_launch_count", "sb200_last_batch_ms", "sb200_set_tiling",
    "sb200_fp64_peak", "sb200_last_export_stats",
    "sb200_group_create", "sb200_group_destroy", "sb200_group_last_error", "sb200_group_size", "sb200_group_member",
    "sb200_group_set_weight", "sb200_group_objective_batch",
)
_NO_STATUS = ("sb200_destroy", "sb200_last_error", "sb200_group_destroy", "sb200_group_last_error")


class StencilB200Error(RuntimeError):
    pass


class _GridDesc(ctypes.Structure):
    _fields_ = [("dim", ctypes.c_int32), ("n", ctypes.c_int32 * 3),
                ("cos_kappa", ctypes.POINTER(ctypes.c_double) * 3),
                ("s2", ctypes.POINTER(ctypes.c_double) * 3),
                ("kcomp", ctypes.POINTER(ctypes.c_double) * 3),
                ("Y", ctypes.c_double), ("Z", ctypes.c_double), ("Y2", ctypes.c_double), ("Z2", ctypes.c_double),
                ("device", ctypes.c_int32)]


class _BatchOpts(ctypes.Structure):
    _fields_ = [("x_begin", ctypes.c_int32), ("x_stride", ctypes.c_int32), ("n_planes", ctypes.c_int32),
                ("flags", ctypes.c_int32), ("seg_size", ctypes.c_int64)]


_lib = None
_cuda_touched = False


def cuda_touched():
    """Has this process called into the CUDA runtime yet (through this library, or through an imported torch)?  A
    process that has must not be forked (Optimize starts its SciPy workers by fork only while this is False)."""
    if _cuda_touched:
        return True
    torch = sys.modules.get("torch")
    try:
        return bool(torch is not None and torch.cuda.is_initialized())
    except Exception:
        return True


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
    op = ctypes.POINTER(_BatchOpts)
    lib.sb200_objective_batch_ex.argtypes = [vp, vp, ctypes.c_int64, op, vp, vp, vp, vp]
    lib.sb200_objective_batch_device_ex.argtypes = [vp, vp, ctypes.c_int64, op, vp, vp]
    lib.sb200_combine_slabs_device.argtypes = [vp, vp, ctypes.c_int32, ctypes.c_int64, vp, vp]
    lib.sb200_last_export_stats.argtypes = [vp, dp, dp]
    i32p = ctypes.POINTER(ctypes.c_int32)
    lib.sb200_group_create.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(_GridDesc), i32p, ctypes.c_int32]
    lib.sb200_group_destroy.argtypes = [vp]
    lib.sb200_group_destroy.restype = None
    lib.sb200_group_last_error.argtypes = [vp]
    lib.sb200_group_last_error.restype = ctypes.c_char_p
    lib.sb200_group_size.argtypes = [vp, i32p]
    lib.sb200_group_member.argtypes = [vp, ctypes.c_int32, ctypes.POINTER(vp)]
    lib.sb200_group_set_weight.argtypes = [vp, ctypes.c_int32, dp, ctypes.c_int64]
    lib.sb200_group_objective_batch.argtypes = [vp, vp, ctypes.c_int64, ctypes.c_int32, op, vp, vp, vp, vp, i32p]
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
        if name not in _NO_STATUS:
            getattr(lib, name).restype = ctypes.c_int
    if lib.sb200_abi_version() != ABI_VERSION:
        raise StencilB200Error("libstencil_b200.so has ABI version %d, expected %d; rebuild it "
                               "(python -m optimize_stencil_b200._build)" % (lib.sb200_abi_version(), ABI_VERSION))
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
    global _cuda_touched
    lib = load_library()
    _cuda_touched = True
    n = ctypes.c_int32(0)
    if lib.sb200_device_count(ctypes.byref(n)) != 0:
        return 0
    return n.value


def fp64_peak_tflops(device=0, seconds=0.25):
    global _cuda_touched
    lib = load_library()
    _cuda_touched = True
    out = ctypes.c_double(0.0)
    rc = lib.sb200_fp64_peak(device, seconds, ctypes.byref(out))
    if rc != 0:
        raise StencilB200Error("sb200_fp64_peak failed (%d): %s" % (rc, lib.sb200_last_error(None).decode()))
    return out.value


def _devices_arg(devices):
    """None -> None (single device); 'all' -> every visible device; an iterable -> a list of ordinals."""
    if devices is None:
        return None
    if isinstance(devices, str):
        if devices != "all":
            raise ValueError("devices must be None, 'all' or a sequence of CUDA ordinals")
        return list(range(max(1, device_count())))
    devs = [int(d) for d in devices]
    if not devs:
        raise ValueError("devices is empty")
    return devs


class Context:
    """One device-resident k-grid: an sb200_ctx, or -- with `devices` -- an sb200_group of one sb200_ctx per GPU
    (exports then run on the group's first device, batches are sharded over all of them).  Not thread-safe; one
    per host thread.

    cos_kappa, s2, kcomp: per-axis 1-D tables exactly as the reference's init2 computes them
    (dispersion.py:268-281, 339-357) -- `dim` arrays each.
    """

    def __init__(self, dim, cos_kappa, s2, kcomp, Y=1.0, Z=1.0, device=0, devices=None):
        global _cuda_touched
        self._lib = load_library()
        _cuda_touched = True
        self._h = ctypes.c_void_p(None)
        self._grp = ctypes.c_void_p(None)
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
        devs = _devices_arg(devices)
        self.devices = devs if devs is not None else [int(device)]
        self.last_shard = ("none", 1)
        if devs is not None and len(devs) > 1:
            arr = (ctypes.c_int32 * len(devs))(*devs)
            rc = self._lib.sb200_group_create(ctypes.byref(self._grp), ctypes.byref(d), arr, len(devs))
            if rc != 0:
                self._grp = ctypes.c_void_p(None)
                raise StencilB200Error("sb200_group_create failed (%d): %s"
                                       % (rc, self._lib.sb200_group_last_error(None).decode()))
            rc = self._lib.sb200_group_member(self._grp, 0, ctypes.byref(self._h))
            if rc != 0:
                raise StencilB200Error("sb200_group_member failed (%d)" % rc)
        else:
            d.device = self.devices[0]
            rc = self._lib.sb200_create(ctypes.byref(self._h), ctypes.byref(d))
            if rc != 0:
                self._h = ctypes.c_void_p(None)
                raise StencilB200Error("sb200_create failed (%d): %s" % (rc, self._lib.sb200_last_error(None).decode()))

    # -- plumbing -------------------------------------------------------------------------------
    def _check(self, rc, what):
        if rc != 0:
            raise StencilB200Error("%s failed (%d): %s" % (what, rc, self._lib.sb200_last_error(self._h).decode()))

    def _gcheck(self, rc, what):
        if rc != 0:
            raise StencilB200Error("%s failed (%d): %s" % (what, rc, self._lib.sb200_group_last_error(self._grp).decode()))

    def close(self):
        if getattr(self, "_grp", None) is not None and self._grp.value:
            self._lib.sb200_group_destroy(self._grp)          # owns every member, including self._h
            self._grp = ctypes.c_void_p(None)
            self._h = ctypes.c_void_p(None)
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
    def _set_weight(self, kind, ptr, count):
        if self._grp.value:
            self._gcheck(self._lib.sb200_group_set_weight(self._grp, kind, ptr, count), "sb200_group_set_weight")
        else:
            self._check(self._lib.sb200_set_weight(self._h, kind, ptr, count), "sb200_set_weight")

    def set_weight_unit(self):
        self._set_weight(SB200_W_UNIT, None, 0)

    def set_weight(self, w):
        """w: scalar-free array broadcastable to the grid.  Arrays that do not depend on the first
        axis of a 3-D grid go to the device as an (ny, nz) plane; everything else dense."""
        w = np.asarray(w, dtype=np.float64)
        x_independent = w.ndim == 3 and (w.shape[0] == 1 or (w.shape[0] == self.shape[0] and w.strides[0] == 0))
        if self.dim == 3 and x_independent:
            plane = _f64(np.broadcast_to(w[:1], (1,) + self.shape[1:])[0])
            self._set_weight(SB200_W_PLANE, _dptr(plane), plane.size)
        else:
            dense = _f64(np.broadcast_to(w, self.shape))
            self._set_weight(SB200_W_DENSE, _dptr(dense), dense.size)

    # -- the objective --------------------------------------------------------------------------
    def _opts(self, x_begin, x_end, x_stride, seg_size, scan, fast=False):
        """sb200_batch_opts, or None for the plain whole-grid call."""
        x_end = self.shape[0] if x_end is None else int(x_end)
        x_begin, x_stride = int(x_begin), int(x_stride)
        flags = SB200_F_AUTO_SCAN if scan == "auto" else (SB200_F_SCAN if scan else 0)
        if fast:
            flags |= SB200_F_FAST
        if x_begin == 0 and x_end == self.shape[0] and x_stride == 1 and not seg_size and not flags:
            return None
        if x_end <= x_begin:
            raise ValueError("empty x range [%d, %d)" % (x_begin, x_end))
        o = _BatchOpts()
        o.x_begin, o.x_stride = x_begin, x_stride
        o.n_planes = (x_end - x_begin + x_stride - 1) // x_stride
        o.flags, o.seg_size = flags, int(seg_size or 0)
        return o

    def objective_batch(self, coeffs, x_begin=0, x_end=None, x_stride=1, seg_size=0, scan=False, shard="auto",
                        fast=None):
        """coeffs: (B, ncoef) full coefficient vectors -> (S, Amin, Amax, nan) arrays of length B.

        fast:     True / False / None: the shorter asin series (3.7e-15 instead of 2.3e-16, ~4 % faster).  None
                  decides by size: a segment (or the batch) of >= 2^28 evaluations is a throughput launch.
        x_begin/x_end/x_stride: the visited x-planes range(x_begin, x_end, x_stride) (k-slabs).
        seg_size: B is a concatenation of independent segments of this many rows; every segment's results are
                  bit-identical to a call on that segment alone.
        scan:     True / False / 'auto': the start-grid screening kernel variant (same results).
        shard:    with a device group: 'auto', 'candidates', 'slabs', 'slabs-contiguous' or 'none'."""
        c = _f64(coeffs)
        if c.ndim == 1:
            c = c.reshape(1, -1)
        if c.ndim != 2 or c.shape[1] != self.ncoef:
            raise ValueError("coeffs must have shape (B, %d), got %s" % (self.ncoef, c.shape))
        B = c.shape[0]
        out = np.empty((4, B))                       # S | Amin | Amax | nan (int64 in the same 8-byte slots)
        base, row = out.ctypes.data, 8 * B
        if fast is None:
            planes = -(-((self.shape[0] if x_end is None else x_end) - x_begin) // x_stride)
            per = float(seg_size or B) * planes * float(np.prod(self.shape[1:]))
            fast = per >= FAST_SERIES_EVALS
        o = self._opts(x_begin, x_end, x_stride, seg_size, scan, fast)
        if self._grp.value:
            used = (ctypes.c_int32 * 2)()
            rc = self._lib.sb200_group_objective_batch(self._grp, c.ctypes.data, B, SHARD_MODES[shard],
                                                       ctypes.byref(o) if o is not None else None, base, base + row,
                                                       base + 2 * row, base + 3 * row, used)
            self._gcheck(rc, "sb200_group_objective_batch")
            self.last_shard = ({v: k for k, v in SHARD_MODES.items()}[used[0]], used[1])
        elif o is None:
            rc = self._lib.sb200_objective_batch(self._h, c.ctypes.data, B, 0, self.shape[0], base, base + row,
                                                 base + 2 * row, base + 3 * row)
            self._check(rc, "sb200_objective_batch")
        else:
            rc = self._lib.sb200_objective_batch_ex(self._h, c.ctypes.data, B, ctypes.byref(o), base, base + row,
                                                    base + 2 * row, base + 3 * row)
            self._check(rc, "sb200_objective_batch_ex")
        return out[0], out[1], out[2], out[3].view(np.int64)

    def objective_batch_device(self, d_coeffs_ptr, nbatch, d_out_ptr, x_begin=0, x_end=None, stream=None, x_stride=1,
                               seg_size=0, scan=False, fast=False):
        """Device-pointer variant (torch tensors' .data_ptr()); asynchronous on `stream`, a raw
        cudaStream_t value (0 / None = CUDA's default stream, i.e. torch's default stream).  Always runs on the
        context's (first) device."""
        o = self._opts(x_begin, x_end, x_stride, seg_size, bool(scan), bool(fast))
        if o is None:
            rc = self._lib.sb200_objective_batch_device(self._h, ctypes.c_void_p(d_coeffs_ptr), int(nbatch), 0,
                                                        self.shape[0], ctypes.c_void_p(d_out_ptr),
                                                        ctypes.c_void_p(stream or 0))
        else:
            rc = self._lib.sb200_objective_batch_device_ex(self._h, ctypes.c_void_p(d_coeffs_ptr), int(nbatch),
                                                           ctypes.byref(o), ctypes.c_void_p(d_out_ptr),
                                                           ctypes.c_void_p(stream or 0))
        self._check(rc, "sb200_objective_batch_device")

    def combine_slabs_device(self, d_parts_ptr, world, nbatch, d_out_ptr, stream=None):
        """[world][4][nbatch] gathered slab partials -> [4][nbatch], fixed shard order, on the device."""
        rc = self._lib.sb200_combine_slabs_device(self._h, ctypes.c_void_p(d_parts_ptr), int(world), int(nbatch),
                                                  ctypes.c_void_p(d_out_ptr), ctypes.c_void_p(stream or 0))
        self._check(rc, "sb200_combine_slabs_device")

    def export(self, coeffs, what=SB200_X_OMEGA, x_begin=0, x_end=None, out=None):
        """Dense omega / sqrtarg / sqrtres for one coefficient vector -> (array, Amin, Amax, nan).

        out: an existing C-contiguous float64 array of the right shape to fill instead of a new one.  A fresh
        NumPy array costs one page fault per 4 KiB on first touch (~10 GB/s with all copy threads), a reused one
        is filled at the PCIe rate."""
        c = _f64(coeffs, (self.ncoef,))
        x_end = self.shape[0] if x_end is None else x_end
        shape = (x_end - x_begin,) + self.shape[1:]
        if out is None:
            out = np.empty(shape, dtype=np.float64)
        elif not (isinstance(out, np.ndarray) and out.dtype == np.float64 and out.shape == shape
                  and out.flags.c_contiguous and out.flags.writeable):
            raise ValueError("out must be a writeable C-contiguous float64 array of shape %s" % (shape,))
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

    def _members(self):
        if not self._grp.value:
            return [self._h]
        out = []
        for i in range(len(self.devices)):
            h = ctypes.c_void_p(None)
            self._gcheck(self._lib.sb200_group_member(self._grp, i, ctypes.byref(h)), "sb200_group_member")
            out.append(h)
        return out

    @property
    def launch_count(self):
        """Kernels launched so far (summed over the devices of a group)."""
        total = 0
        for h in self._members():
            n = ctypes.c_int64(0)
            self._check(self._lib.sb200_launch_count(h, ctypes.byref(n)), "sb200_launch_count")
            total += n.value
        return total

    @property
    def last_batch_ms(self):
        ms = ctypes.c_double(0.0)
        self._check(self._lib.sb200_last_batch_ms(self._h, ctypes.byref(ms)), "sb200_last_batch_ms")
        return ms.value

    @property
    def last_export_stats(self):
        """(bytes delivered to host buffers, wall seconds) of the most recent export / group_velocity call."""
        b, t = ctypes.c_double(0.0), ctypes.c_double(0.0)
        self._check(self._lib.sb200_last_export_stats(self._h, ctypes.byref(b), ctypes.byref(t)), "sb200_last_export_stats")
        return b.value, t.value
