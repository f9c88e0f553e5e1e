"""Where a cold process spends its first second: import, first CUDA call (context), first launch (module load)."""
import time
t0 = time.perf_counter()
import os, sys  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
t1 = time.perf_counter()
import optimize_stencil_b200 as sb  # noqa: E402
from optimize_stencil_b200 import _capi  # noqa: E402
t2 = time.perf_counter()
lib = _capi.load_library()
t3 = time.perf_counter()
n = sb.device_count()
t4 = time.perf_counter()
N = 128
x = np.linspace(0, np.pi, N)
ctx = sb.Context(3, [np.cos(x)] * 3, [0.5 * (1 - np.cos(x))] * 3, [x] * 3)
t5 = time.perf_counter()
co = np.zeros((4096, 13)); co[:, 0] = 0.5; co[:, 1:4] = 1.0
ctx.objective_batch(co[:3])
t6 = time.perf_counter()
ctx.objective_batch(co[:3])
t7 = time.perf_counter()
ctx.objective_batch(co, scan=True)
t8 = time.perf_counter()
ctx.objective_batch(co, scan=True)
t9 = time.perf_counter()
print("numpy import %.3f | package import %.3f | dlopen %.3f | device_count (cuInit) %.3f | context create %.3f | "
      "first launch (3 rows, precise) %.3f | second %.4f | first screening launch (4096 rows) %.3f | second %.4f  [s]"
      % (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t7 - t6, t8 - t7, t9 - t8))
