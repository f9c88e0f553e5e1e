"""Weight functions for the objective (drop-in for the reference's extended_stencil/weight.py).

`weight_functions[name](*params)` returns a setter that assigns `dispersion.w`; the Dispersion
object uploads it to the GPU before its next evaluation.  The values are the reference's
(weight.py:4-31); the arrays are built without touching the dense k array:

  xaxis       1 on the kx axis, 0 elsewhere.  The reference allocates a dense zeros_like(k); here it
              is a read-only broadcast view of one (ky[,kz]) plane with the same shape and values,
              so a 512^3 grid costs 2 MiB instead of 1 GiB and goes to the device as a plane table.
  xaxis_soft  exp(-(ky/width)^2 [-(kz/width)^2]) with the reference's shapes (1,N) / (1,N,N).
  equal       is the absence of a factory: `w = 1.0` (dispersion.py:40).
"""
import numpy as np


def weight_xaxis():
    def set_weight_xaxis(dispersion):
        if dispersion.dim not in (2, 3):
            raise ValueError('Dispersion object not 2D or 3D')
        N = dispersion.N
        plane = np.zeros((1,) + (N,) * (dispersion.dim - 1))
        plane[(0,) * dispersion.dim] = 1.0
        dispersion.w = np.broadcast_to(plane, (N,) * dispersion.dim)
    return set_weight_xaxis


def weight_xaxis_soft(width=0.1):
    def set_weight_xaxis_soft(dispersion):
        if dispersion.dim == 2:
            dispersion.w = np.exp(-(dispersion.ky / width) ** 2)
        elif dispersion.dim == 3:
            dispersion.w = np.exp(-(dispersion.ky / width) ** 2 - (dispersion.kz / width) ** 2)
        else:
            raise ValueError('Dispersion object not 2D or 3D')
    return set_weight_xaxis_soft


weight_functions = dict(xaxis=weight_xaxis, xaxis_soft=weight_xaxis_soft)
