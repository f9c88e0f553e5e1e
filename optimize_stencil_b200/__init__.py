"""optimize_stencil_b200 -- B200-native objective for Ablinne/optimize-stencil.

Layout (only what the hot path needs):
  csrc/               hand-written sm_100a CUDA kernels + the C-ABI (libstencil_b200.so)
  _capi.py            ctypes binding of include/stencil_b200.h
  extended_stencil/   drop-in mirror of the reference's Python API (Stencil*, Dispersion2D/3D,
                      weight_functions, Optimize) whose grid arithmetic runs on the GPU
  sharding.py         candidate / k-slab sharding over torch.distributed (one process per GPU)
"""
from ._capi import Context, StencilB200Error, device_count, fp64_peak_tflops, load_library  # noqa: F401

__all__ = ["Context", "StencilB200Error", "device_count", "fp64_peak_tflops", "load_library"]
