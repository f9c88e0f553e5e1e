"""Debug aid: run one Optimize case sequentially and log every evaluated row (coefficients and results)."""
import os, sys, io, contextlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from optimize_stencil_b200 import _capi
if os.environ.get("ORACLE_BACKED"):
    import fake_context
    _capi.Context = fake_context.OracleContext
from optimize_stencil_b200 import extended_stencil as es
from test_dropin_api import default_args, OPT_RUNS
Ctx = _capi.Context
orig = Ctx.objective_batch
rows = []
def logged(self, coeffs, *a, **k):
    r = orig(self, coeffs, *a, **k)
    c = np.atleast_2d(coeffs)
    for b in range(len(c)):
        rows.append(np.concatenate([c[b], [r[0][b], r[1][b], r[2][b], float(r[3][b] > 0), self.shape[0]]]))
    return r
Ctx.objective_batch = logged
es.Optimize.lockstep = False
with contextlib.redirect_stdout(io.StringIO()):
    opt = es.Optimize(default_args(**OPT_RUNS[sys.argv[1]]))
    x, f = opt.optimize()
np.save(sys.argv[2], np.array(rows))
print("result", x, f, "rows", len(rows))
