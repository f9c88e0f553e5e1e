#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
{
echo "== warp tiles"; python tools/export_bench.py 512 2>&1 | grep -E "^export|Context.export" | head -5
echo "== CTA tiles (round 1)"; SB200_EXPORT_CTA_TILES=1 python tools/export_bench.py 512 2>&1 | grep -E "^export omega"
echo "== 256^3"; python tools/export_bench.py 256 2>&1 | grep -E "^export omega"; SB200_EXPORT_CTA_TILES=1 python tools/export_bench.py 256 2>&1 | grep -E "^export omega"
python - <<'PY'
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import np_oracle as O, optimize_stencil_b200 as sb, os
for dim, N, Y, Z in ((3, 77, 1.3, 0.8), (3, 130, 1.0, 1.0), (2, 200, 1.0, 1.0), (2, 33, 2.0, 1.0)):
    grid = O.Grid(dim, N, Y, Z); cos, s2, kc = grid.flat_tables()
    fam = "StencilFree3D" if dim == 3 else "StencilFree2D"
    rng = np.random.default_rng(N)
    with sb.Context(dim, cos, s2, kc, Y=Y, Z=Z) as ctx:
        for dt in (0.3, 0.55):
            p = 0.02 * rng.standard_normal(len(O.parameter_names(fam))); p[0] = dt
            co = O.coefficients(fam, p)
            om, lo, hi, nn = ctx.export(co, 0)
            os.environ["SB200_EXPORT_CTA_TILES"] = "1"
            with sb.Context(dim, cos, s2, kc, Y=Y, Z=Z) as old:
                om0, lo0, hi0, nn0 = old.export(co, 0)
            del os.environ["SB200_EXPORT_CTA_TILES"]
            same = np.array_equal(om.view(np.int64), om0.view(np.int64)) and (lo, hi, nn) == (lo0, hi0, nn0)
            print(dim, N, dt, "bitwise equal to the CTA-tiled kernel:", same, "nan", nn)
PY
} > gpurun_out/r2x_export.log 2>&1
cat gpurun_out/r2x_export.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_cli.py -m gpu -q --timeout 600 2>&1 | tail -3
