"""Dense omega export (sb200::export_kernel): device-side throughput against the HBM-write roofline.

    python tools/export_bench.py [N]     (3-D N^3 grid, Lehe dt = 0.96; default N = 512 -> 1 GiB of omega)
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import optimize_stencil_b200 as sb  # noqa: E402
from optimize_stencil_b200.extended_stencil import StencilSymmetric3D  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
x = np.linspace(0, np.pi, N)
cos, s2 = np.cos(x), 0.5 * (1. - np.cos(x))
ctx = sb.Context(3, [cos] * 3, [s2] * 3, [x] * 3)
dlx = 0.25 * (1 - 1 / 0.96 ** 2 * np.sin(np.pi * 0.96 / 2.0))
co = StencilSymmetric3D().coefficients_batch([0.96, .125, .125, .125, dlx, 0, 0])[0]
out = torch.empty(N ** 3, dtype=torch.float64, device="cuda")
stats = torch.empty(3, dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
hbm = peaks.get("hbm_gbs", 6650.0)
for what, name in ((0, "omega"), (1, "sqrtarg"), (2, "sqrtres")):
    for _ in range(3):
        ctx.export_device(co, out.data_ptr(), what, d_stats_ptr=stats.data_ptr(), stream=st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 10
    for _ in range(K):
        ctx.export_device(co, out.data_ptr(), what, d_stats_ptr=stats.data_ptr(), stream=st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    gbs = N ** 3 * 8 / (ms * 1e-3) / 1e9
    print("export %-8s %d^3: %.3f ms  %.1f GB/s written (algorithmic 8 B/point)  %.2f of %s HBM copy peak %.0f GB/s  "
          "%.1f G points/s" % (name, N, ms, gbs, gbs / hbm, "measured" if peaks else "fallback", hbm, N ** 3 / ms / 1e6))

# End to end: Context.export into a fresh NumPy array (kernel + chunked, double-buffered pinned D2H + the host-side
# copy into the caller's pageable buffer); the rate is what sb200_last_export_stats reports for the call.
import time  # noqa: E402
for rep in range(4):
    t0 = time.perf_counter()
    om = ctx.export(co, 0)[0]
    wall = time.perf_counter() - t0
    nbytes, secs = ctx.last_export_stats
    print("Context.export(omega) %d^3 -> NumPy: %.1f ms wall, %.1f ms inside sb200_export = %.1f GB/s delivered "
          "(copy threads: %s)" % (N, 1e3 * wall, 1e3 * secs, nbytes / secs / 1e9, os.environ.get("SB200_COPY_THREADS", "default")))
    del om
pinned = torch.empty(N ** 3, dtype=torch.float64).pin_memory()
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pinned.copy_(out, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
print("plain D2H of the same field into pinned memory: %.1f ms = %.1f GB/s (the PCIe rate the export can approach)"
      % (1e3 * dt, N ** 3 * 8 / dt / 1e9))
dst = np.empty(N ** 3)
t0 = time.perf_counter()
np.copyto(dst, pinned.numpy())
dt = time.perf_counter() - t0
print("one host thread copying it from pinned into a fresh NumPy array: %.1f ms = %.1f GB/s" % (1e3 * dt, N ** 3 * 8 / dt / 1e9))
