#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric on its config: k-point x candidate objective evals/s (fp64).

Workload (config.workload): BASELINE.json configs[3] -- "batched objective: 4096 random coefficient
vectors x 512^3 k-grid, candidate-sharded across 1/2/4/8 GPUs" (the configuration the metric's
"1/2/4/8 GPU" is quoted on; it fits one GPU: the kernel needs O(N) tables, no N^3 array).  One STEP =
one pass of the hot path over the whole batch: 4096 x 512^3 = 5.5e11 objective evaluations, each
yielding (sum w(omega-k)^2, min sqrtarg, max sqrtarg) contributions.  Total work is fixed as N grows
("scaling": "strong"); rank g evaluates candidates [g*B/G, (g+1)*B/G) and ONE all-gather of the
(4 x B/G) result blocks combines them.

  python bench.py [--gpus N] [--steps K] [--warmup W]                   (N > 1: launched by torchrun)
  python bench.py --impl reference ...   the reference's CPU path (NumPy port in a multiprocessing.Pool)

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (inputs already in HBM);
`e2e` = the same through the host-buffer API (pinned H2D of the coefficients, kernels, collective,
D2H of the results inside the timed region); `roofline` = FP64-pipe roofline of the objective kernel
with F_alg = 61 flop per evaluation; `cpu_baseline` = oracle port on this box's cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F_ALG_3D = 61.0            # flop per k-point x candidate
FP64_NOMINAL_TFLOPS = 37.2  # 148 SM x 64 lanes x 2 x 1.965 GHz
METRIC = "k-point*candidate objective evals/s (fp64)"
UNIT = "evals/s"
XC = np.array([0.5, .02, .03, .01, .02, .03, .01, -.02, -.03, -.01])


def workload_parameters(B, seed=0):
    """B StencilFree3D parameter vectors: x_c + 1e-3 N(0,1), numpy default_rng(seed)."""
    return XC + 1e-3 * np.random.default_rng(seed).standard_normal((B, 10))


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smmax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for line in f:
                t = [v.strip() for v in line.split(",")]
                if len(t) < 8:
                    continue
                try:
                    sm.append(float(t[0])); smmax.append(float(t[1])); power.append(float(t[2]))
                except ValueError:
                    continue
                for name, v in zip(names, t[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smmax)), power_w_max=float(max(power)),
                    samples=len(sm), reasons=sorted(reasons))


# ---------------------------------------------------------------------------------------------------
# CPU arm: the reference's multiprocessing path, restated by the oracle (oracle/np_oracle.py)
# ---------------------------------------------------------------------------------------------------
def _cpu_worker(task):
    """One candidate over the full N^3 grid, evaluated x-slab by x-slab with the NumPy port of
    Dispersion3D.norm (the reference needs ~7 N^3-sized temporaries; slabs keep that at ~1 GiB/worker)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import np_oracle as O
    coeffs, N, slab = task
    S = 0.0
    for a in range(0, N, slab):
        S += O.device_contract(O.Grid(3, N, xslab=(a, min(a + slab, N))), coeffs)[0]
    return S * O.norm_scale(3, N)


def cpu_pool_size():
    """The reference's own choice (optimize.py:173-179): physical cores, +1 if SMT; capped by memory."""
    try:
        import psutil
        n = psutil.cpu_count(logical=False) or os.cpu_count()
        if (psutil.cpu_count(logical=True) or n) > n:
            n += 1
        avail = psutil.virtual_memory().available
        n = max(1, min(n, int(avail // (3 << 30))))
    except ImportError:
        n = os.cpu_count() or 1
    return n


def cpu_reference_steps(N, steps, warmup, seed=0):
    """Times `steps` samples of the workload on the host cores; returns (evals/s, cores, sample text, ms/step)."""
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import np_oracle as O
    P = cpu_pool_size()
    per_step = P  # one candidate per worker and step: about 6-10 s of all-core work at 512^3
    params = workload_parameters(per_step * (steps + warmup), seed)
    coeffs = [O.coefficients("StencilFree3D", p) for p in params]
    slab = max(1, min(N, (1 << 24) // (N * N)))  # ~16.8 M points per slab = the reference at 256^3
    times = []
    with mp.Pool(P) as pool:
        for s in range(steps + warmup):
            tasks = [(c, N, slab) for c in coeffs[s * per_step:(s + 1) * per_step]]
            t0 = time.perf_counter()
            out = pool.map(_cpu_worker, tasks, chunksize=1)
            dt = time.perf_counter() - t0
            assert np.all(np.isfinite(out))
            if s >= warmup:
                times.append(dt)
    total = sum(times)
    evals = per_step * steps * float(N) ** 3
    sample = ("%d of the 4096 candidates per step x full %d^3 grid, each evaluated in x-slabs of %d planes by the "
              "NumPy port of Dispersion3D.norm (oracle/np_oracle.py) in a multiprocessing.Pool(%d)"
              % (per_step, N, slab, P))
    return evals / total, P, sample, 1e3 * total / steps


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    value, cores, sample, ms = cpu_reference_steps(a.grid, a.steps, a.warmup)
    line = dict(metric=METRIC, value=value, unit=UNIT, impl="reference", n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
                ms_per_step=ms, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64",
                data="synthetic", config=config_dict(a, extra={"cpu_sample_per_step": sample}),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))
    return 0


def config_dict(a, extra=None):
    c = {"workload": "BASELINE.json configs[3]: batched objective, %d random StencilFree3D coefficient vectors x "
                     "%d^3 k-grid, weight=equal, Y=Z=1, candidate-sharded" % (a.batch, a.grid),
         "candidates": a.batch, "grid": [a.grid] * 3, "evals_per_step": a.batch * a.grid ** 3,
         "parameters": "x_c + 1e-3*N(0,1), numpy default_rng(0), x_c per SURVEY.md 8d",
         "parallelism": "candidates/%d" % a.gpus,
         "l2": "no L2 flush needed: the kernel reads O(N) tables and 13 doubles per candidate; there is no "
               "N^3 input to keep warm (algorithmic HBM traffic ~ 0)"}
    if extra:
        c.update(extra)
    return c


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def run_gpu_arm(a):
    import torch
    import torch.distributed as dist
    import optimize_stencil_b200 as sb
    from optimize_stencil_b200 import sharding
    from optimize_stencil_b200.extended_stencil import StencilFree3D, Dispersion3D

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the objective has no CPU fallback (use --impl reference "
                         "for the CPU arm)")
    if world != a.gpus:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d (launch N>1 with torch.distributed.run)" % (a.gpus, world))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    # Keep stdout to the one JSON line: anything libraries print while the communicator comes up (NCCL's
    # version banner, debug output) is sent to stderr until the warm-up is over.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    N, B = a.grid, a.batch
    stencil = StencilFree3D()
    disp = Dispersion3D(1.0, 1.0, N, stencil)   # the drop-in object: NumPy tables exactly as the reference
    disp.device = local
    ctx = disp._context()
    params = workload_parameters(B)
    coeffs_host = stencil.coefficients_batch(params)
    coeffs_dev = torch.from_numpy(coeffs_host).to(device)
    sharded = sharding.ShardedObjective(sharding.device_evaluator(ctx), nx=N, mode="candidates")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    peak = sb.fp64_peak_tflops(local, 0.4)   # measured DFMA roofline denominator, same box, same run

    # ---- device-resident timing: K steps, CUDA events on the launching (torch current) stream -------
    for _ in range(a.warmup):
        out = sharded.evaluate(coeffs_dev)
    barrier()
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    lo, hi = sharded.my_candidates(B)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    evaluate_local = sharded.evaluate_local

    def timed_local(block, x0, x1, _k=[0]):
        kev[_k[0]][0].record()
        r = evaluate_local(block, x0, x1)
        kev[_k[0]][1].record()
        _k[0] += 1
        return r

    sharded.evaluate_local = timed_local
    barrier()
    t_wall0 = time.perf_counter()
    for s in range(a.steps):
        ev[s][0].record()
        out = sharded.evaluate(coeffs_dev)
        ev[s][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    sharded.evaluate_local = evaluate_local
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launch_count - launches0
    dev_ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    kern_ms = sum(e0.elapsed_time(e1) for e0, e1 in kev)
    t = torch.tensor([dev_ms, t_wall * 1e3], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max = t.tolist()
    evals_per_step = float(B) * float(N) ** 3
    value = evals_per_step * a.steps / (dev_ms_max * 1e-3)

    # ---- sanity: the result is finite, gate-clean, and identical on every rank -----------------------
    res = out.cpu().numpy()
    assert res.shape == (4, B) and np.all(np.isfinite(res[0])) and np.all(res[3] == 0), "bench result is not clean"
    chk = torch.tensor([float(np.sum(res[0]))], dtype=torch.float64, device=device)
    if world > 1:
        chk2 = chk.clone()
        dist.all_reduce(chk2, op=dist.ReduceOp.MAX)
        assert chk2.item() == chk.item(), "ranks disagree on the gathered result"

    # ---- end to end through the host-buffer API -------------------------------------------------------
    for _ in range(max(1, min(a.warmup, 2))):
        sharding.host_evaluate(sharded, coeffs_host, device)
    barrier()
    t0 = time.perf_counter()
    for s in range(a.steps):
        S, Amin, Amax, nan = sharding.host_evaluate(sharded, coeffs_host, device)
    barrier()
    e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = evals_per_step * a.steps / e2e_t.item()
    assert np.array_equal(S, res[0]), "host-buffer path and device-resident path disagree"
    norms = disp.norm_batch(params[:4]) if rank == 0 else None   # the user-facing call, spot check
    if rank == 0:
        scale = (np.pi / (N - 1)) ** 3
        assert np.allclose(norms, res[0][:4] * scale, rtol=1e-12, atol=0), "Dispersion3D.norm_batch disagrees"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant (only) kernel: per launch, measured live ------------------------------
    launch_evals = float(hi - lo) * float(N) ** 3
    launch_ms = kern_ms / a.steps
    achieved = launch_evals * F_ALG_3D / (launch_ms * 1e-3) / 1e12
    traffic = inst = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            prof = json.load(open(tpath))
            traffic = prof.get("objective_kernel_dram_bytes_per_launch")
            inst = prof.get("objective_kernel_fp64_inst_per_eval")
        except Exception:
            traffic = inst = None
    roofline = dict(bound="fp64", achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak,
                    traffic=traffic, kernel="sb200::objective_kernel_fa<4,true,0>",
                    flop_per_eval=F_ALG_3D, evals_per_launch=launch_evals, launch_ms=launch_ms,
                    peak_source="DFMA microbenchmark (sb200_fp64_peak) measured in this run on this GPU; "
                                "MEASURED_PEAKS.json has no FP64 entry",
                    peak_nominal=FP64_NOMINAL_TFLOPS, frac_of_nominal=achieved / FP64_NOMINAL_TFLOPS,
                    note="FP64-pipe bound, not HBM or tensor: inputs are O(N) tables + 13 doubles per candidate. "
                         "`achieved` prices every evaluation at the reference formulation's 61 flop (SURVEY.md 8d); "
                         "the kernel regroups sqrtarg and evaluates asin without a square root in the common class, "
                         "so it EXECUTES far fewer FP64 instructions than that (see `executed`) and frac can exceed 1")
    if inst:
        # what the FP64 pipe really does: executed thread-instructions per evaluation (ncu, profiles/) x the
        # evaluation rate of this run, against the measured DFMA issue rate (peak TFLOP/s / 2 flop per DFMA)
        rate = launch_evals / (launch_ms * 1e-3) * inst
        roofline["executed"] = dict(fp64_inst_per_eval=inst, fp64_inst_per_s=rate,
                                    pipe_frac_of_measured=rate / (peak * 1e12 / 2.0),
                                    source="profiles/traffic.json (ncu inst_executed by opcode of the bench's own launch)")

    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        v, cores, sample, _ = cpu_reference_steps(N, 1, 0, seed=0)
        cpu = dict(value=v, unit=UNIT, cores=cores, kind="port", sample=sample)

    h2d = (hi - lo) * 13 * 8
    d2h = 4 * B * 8
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=a.steps, warmup=a.warmup,
                ms_per_step=dev_ms_max / a.steps, higher_is_better=True, scaling="strong", vs_baseline=None,
                dtype="f64", data="synthetic", config=config_dict(a), clocks=clocks,
                e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                         api="sharding.host_evaluate -> Context.objective_batch_device (ctypes C-ABI)"),
                gpu_launches=int(launches) * world, roofline=roofline, cpu_baseline=cpu,
                wall_ms_per_step=wall_ms_max / a.steps,
                fp64_fraction=dict(of_measured_peak=value / world * F_ALG_3D / (peak * 1e12),
                                   of_nominal_peak=value / world * F_ALG_3D / (FP64_NOMINAL_TFLOPS * 1e12)))
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--batch", type=int, default=4096, help="candidates per step (default: the named config)")
    ap.add_argument("--grid", type=int, default=512, help="k-grid points per axis (default: the named config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    if a.warmup < 3 and a.impl == "b200":
        a.warmup = 3
    return run_reference_arm(a) if a.impl == "reference" else run_gpu_arm(a)


if __name__ == "__main__":
    sys.exit(main())
