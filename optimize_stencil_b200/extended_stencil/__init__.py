"""Drop-in replacement for the reference's `extended_stencil` package, backed by the sm_100a library.

    from optimize_stencil_b200 import extended_stencil            # or alias it:
    import sys; sys.modules['extended_stencil'] = extended_stencil

Exports the same names as the reference's extended_stencil/__init__.py:23-33.
"""
import numpy as np


def minmax(arr):
    return np.min(arr), np.max(arr)


from .stencil import *            # noqa: F401,F403,E402
from .stencil import namedarray, StencilFlags, Stencil, get_stencil  # noqa: F401,E402
from .dispersion import Dispersion, Dispersion2D, Dispersion3D       # noqa: F401,E402
from .optimize import Optimize, make_single_max_constraint, make_single_min_constraint  # noqa: F401,E402
from .weight import weight_functions                                  # noqa: F401,E402

__version__ = "0.1.0+b200"
