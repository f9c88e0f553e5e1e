"""The C-ABI library loads on a machine without a GPU and exports every symbol include/stencil_b200.h
declares; compute entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "stencil_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sb200_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from optimize_stencil_b200 import _capi
    assert sorted(_capi.EXPORTED_SYMBOLS) == header_symbols()


def test_library_exports_every_declared_symbol():
    from optimize_stencil_b200 import _build, _capi
    _build.build()  # no-op when up to date; nvcc cross-compiles without a GPU
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for name in header_symbols():
        assert getattr(lib, name) is not None, name
    assert _capi.load_library().sb200_abi_version() == _capi.ABI_VERSION == 2


def test_cuobjdump_shows_sm100a_only():
    import shutil
    import subprocess
    from optimize_stencil_b200 import _capi
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([tool, "-lelf", _capi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    import optimize_stencil_b200 as sb
    from optimize_stencil_b200 import extended_stencil as es
    assert sb.device_count() == 0
    with pytest.raises(sb.StencilB200Error, match="no CPU fallback"):
        sb.Context(2, [np.ones(4)] * 2, [np.zeros(4)] * 2, [np.zeros(4)] * 2)
    d = es.Dispersion2D(1, 10, es.StencilFree2D())
    with pytest.raises(sb.StencilB200Error):
        d.norm([0.3, 0, 0, 0, 0])
    with pytest.raises(sb.StencilB200Error):
        sb.fp64_peak_tflops()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under optimize_stencil_b200/ may reference it."""
    pkg = os.path.join(ROOT, "optimize_stencil_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "np_oracle" not in text and "import oracle" not in text and "ref_shim" not in text, f
