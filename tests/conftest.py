import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class GoldenCases:
    """tests/golden/objective_cases.npz (written by oracle/make_golden.py from the unmodified reference)."""

    def __init__(self):
        z = np.load(os.path.join(GOLDEN, "objective_cases.npz"))
        self.z = z
        self.n = len(z["stencil"])
        self.grids = []
        for gi in range(int(z["ngrids"])):
            self.grids.append(dict(key=tuple(z["grid%d_key" % gi]), cos=z["grid%d_cos" % gi],
                                   s2=z["grid%d_s2" % gi], kcomp=z["grid%d_kcomp" % gi]))

    def case(self, i):
        z = self.z
        c = {k: z[k][i] for k in ("stencil", "dim", "N", "Y", "Z", "wkind", "wparam", "dtmul", "nparams",
                                  "tag", "stencil_ok", "dt_ok", "norm", "gated", "Amin", "Amax", "S", "grid_id")}
        c["stencil"], c["wkind"], c["tag"] = str(c["stencil"]), str(c["wkind"]), str(c["tag"])
        c["dim"], c["N"], c["nparams"] = int(c["dim"]), int(c["N"]), int(c["nparams"])
        c["Y"], c["Z"], c["dtmul"] = float(c["Y"]), float(c["Z"]), float(c["dtmul"])
        c["params"] = z["params"][i][:c["nparams"]]
        c["coeffs"] = z["coeffs"][i][:(7 if c["dim"] == 2 else 13)]
        c["wparams"] = () if np.isnan(c["wparam"]) else (float(c["wparam"]),)
        c["omega_samples"] = z["omega_samples"][i]
        c["grid"] = self.grids[int(c["grid_id"])]
        c["id"] = i
        key = "omega_full_%d" % i
        c["omega_full"] = z[key] if key in z.files else None
        return c

    def sample_indices(self, i, size):
        rng = np.random.default_rng(1000 + i)
        idx = rng.integers(0, size, 96)
        idx[0], idx[1] = 0, size - 1
        return idx


@pytest.fixture(scope="session")
def golden():
    return GoldenCases()


def relerr(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    with np.errstate(invalid="ignore", divide="ignore"):
        e = np.abs(a - b) / np.abs(b)
    return np.where(b == 0, np.abs(a), e)
