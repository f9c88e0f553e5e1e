"""TEST INFRASTRUCTURE ONLY -- loads the UNMODIFIED reference from /root/reference or baseline/_ref.

The reference (Ablinne/optimize-stencil) is pure Python but no longer imports on
NumPy >= 2 (it uses np.asfarray / np.asscalar, removed upstream) -- SURVEY.md
section 8(c).  This module installs the three aliases from *outside* the
reference tree and imports it.  It is used by oracle/make_golden.py (to produce
the committed fixtures under tests/golden/) and by tests that are skipped when
/root/reference is absent (it does not exist on the GPU box).

Nothing in the product package (optimize_stencil_b200/) imports this file.
"""
import os
import sys
import importlib

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root():
    """/root/reference (this container) or baseline/_ref (oracle/install_reference.sh: the same files, pip-installed
    where they can travel to the GPU box)."""
    cands = [os.environ.get("OPTSTENCIL_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "extended_stencil")):
            return c
    return cands[1]


REFERENCE_ROOT = _find_root()


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "extended_stencil"))


def load_reference():
    """Import and return the reference's `extended_stencil` package (shimmed)."""
    import numpy as np

    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if not hasattr(np, "asfarray"):
        np.asfarray = lambda a, dtype=float: np.asarray(a, dtype=dtype)
    if not hasattr(np, "asscalar"):
        np.asscalar = lambda a: np.asarray(a).item()
    sys.dont_write_bytecode = True  # the reference tree is read-only
    # The product ships a drop-in that can also be aliased as `extended_stencil`;
    # make sure we really get the reference's files here.
    for name in [m for m in sys.modules if m == "extended_stencil" or m.startswith("extended_stencil.")]:
        del sys.modules[name]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        mod = importlib.import_module("extended_stencil")
    finally:
        sys.path.remove(REFERENCE_ROOT)
    assert os.path.realpath(mod.__file__).startswith(os.path.realpath(REFERENCE_ROOT)), mod.__file__
    return mod
