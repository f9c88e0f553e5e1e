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
  python bench.py --impl reference ...   the reference's CPU path: the UNMODIFIED reference classes from
                                         baseline/_ref (oracle/install_reference.sh) in a multiprocessing.Pool,
                                         else the NumPy port (oracle/np_oracle.py)

Prints ONE JSON line (rank 0).
  value        device-resident throughput (inputs already in HBM), CUDA events on the launching stream, max over ranks
  e2e          the same through the user-facing call, Dispersion3D.norm_batch(params) (coefficient fill on the host,
               pinned H2D, kernel, D2H, the reference's gate; N > 1: every rank its shard, then one all-gather of the
               norms), host wall clock between barriers
  roofline     FP64-pipe roofline of the objective kernel, F_alg = 61 flop per evaluation, against the
               DFMA peak measured in this run; `executed` = what the pipe really issues (ncu instruction counts)
  parity       rows of THIS run's batch re-evaluated by the plain-C oracle outside the timed region
  cpu_baseline the reference's CPU path on this box's cores, bounded sample
  extra        the other named configs of BASELINE.json (1, 2, 3, 5), each with its own timing and oracle check
"""
import argparse
import contextlib
import io
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
F_ALG_2D = 57.0
FP64_NOMINAL_TFLOPS = 37.2  # 148 SM x 64 lanes x 2 x 1.965 GHz
METRIC = "k-point*candidate objective evals/s (fp64)"
UNIT = "evals/s"
XC = np.array([0.5, .02, .03, .01, .02, .03, .01, -.02, -.03, -.01])


def workload_parameters(B, seed=0):
    """B StencilFree3D parameter vectors: x_c + 1e-3 N(0,1), numpy default_rng(seed)."""
    return XC + 1e-3 * np.random.default_rng(seed).standard_normal((B, 10))


def oracle_path():
    p = os.path.join(ROOT, "oracle")
    if p not in sys.path:
        sys.path.insert(0, p)


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
# CPU arm: the reference's multiprocessing path (optimize.py:169-190: one Dispersion3D.norm sweep per task)
# ---------------------------------------------------------------------------------------------------
def reference_kind():
    """'reference' when the unmodified reference is importable on this box (baseline/_ref or /root/reference),
    else 'port' (oracle/np_oracle.py, the pinned NumPy restatement)."""
    oracle_path()
    import ref_shim
    return "reference" if ref_shim.reference_available() else "port"


def _cpu_worker(task):
    """One candidate over the full N^3 grid, x-slab by x-slab (the reference needs ~7 N^3-sized temporaries; slabs
    of 2^24 points keep a worker at ~1 GiB and are the size the reference was measured at in SURVEY.md: 256^3)."""
    oracle_path()
    kind, params, N, slab = task
    if kind == "reference":
        import ref_slab
        return ref_slab.norm_slabwise(params, N, slab)
    import np_oracle as O
    coeffs = O.coefficients("StencilFree3D", params)
    S = 0.0
    for a in range(0, N, slab):
        S += O.device_contract(O.Grid(3, N, xslab=(a, min(a + slab, N))), coeffs)[0]
    return S * O.norm_scale(3, N)


def _cpu_warm(task):
    """Build the per-process tables (the reference's init2 state) before anything is timed."""
    oracle_path()
    kind, N, slab = task
    if kind == "reference":
        import ref_slab
        ref_slab.reference_state(N, slab)
    return os.getpid()


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
    """Times `steps` samples of the workload on the host cores -> (evals/s, cores, kind, sample text, ms/step)."""
    import multiprocessing as mp
    kind = reference_kind()
    P = cpu_pool_size()
    per_step = P  # one candidate per worker and step: about 6-10 s of all-core work at 512^3
    params = workload_parameters(per_step * (steps + warmup), seed)
    slab = max(1, min(N, (1 << 24) // (N * N)))  # ~16.8 M points per slab = the reference at 256^3
    times = []
    with mp.get_context("spawn").Pool(P) as pool:
        pool.map(_cpu_warm, [(kind, N, slab)] * (4 * P), chunksize=1)
        for s in range(steps + warmup):
            tasks = [(kind, p, N, slab) for p in params[s * per_step:(s + 1) * per_step]]
            t0 = time.perf_counter()
            out = pool.map(_cpu_worker, tasks, chunksize=1)
            dt = time.perf_counter() - t0
            assert np.all(np.isfinite(out)) and np.all(np.asarray(out) < 1e5), "CPU arm: a candidate was gated"
            if s >= warmup:
                times.append(dt)
    total = sum(times)
    evals = per_step * steps * float(N) ** 3
    what = ("the UNMODIFIED reference's Dispersion3D.norm (baseline/_ref, x-axis tables overridden per slab: "
            "oracle/ref_slab.py)" if kind == "reference" else
            "the NumPy port of Dispersion3D.norm (oracle/np_oracle.py; the unmodified reference is not installed "
            "on this box)")
    sample = ("%d of the 4096 candidates per step x full %d^3 grid, each evaluated in x-slabs of %d planes by %s "
              "in a multiprocessing.Pool(%d)" % (per_step, N, slab, what, P))
    return evals / total, P, kind, sample, 1e3 * total / steps


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    value, cores, kind, sample, ms = cpu_reference_steps(a.grid, a.steps, a.warmup)
    line = dict(metric=METRIC, value=value, unit=UNIT, impl="reference", n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
                ms_per_step=ms, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64",
                data="synthetic", config=config_dict(a),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind=kind, sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))
    return 0


def config_dict(a):
    return {"workload": "BASELINE.json configs[3]: batched objective, %d random StencilFree3D coefficient vectors x "
                        "%d^3 k-grid, weight=equal, Y=Z=1, candidate-sharded" % (a.batch, a.grid),
            "candidates": a.batch, "grid": [a.grid] * 3, "evals_per_step": a.batch * a.grid ** 3,
            "parameters": "x_c + 1e-3*N(0,1), numpy default_rng(0), x_c per SURVEY.md 8d",
            "parallelism": "candidates/%d" % a.gpus,
            "l2": "no L2 flush needed: the kernel reads O(N) tables and 13 doubles per candidate; there is no "
                  "N^3 input to keep warm (algorithmic HBM traffic ~ 0)"}


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
class Env:
    """Process-group plumbing shared by the main record and the extras."""

    def __init__(self, a):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the objective has no CPU fallback (use --impl reference "
                             "for the CPU arm)")
        if self.world != a.gpus:
            raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d (launch N>1 with torch.distributed.run)"
                             % (a.gpus, self.world))
        torch.cuda.set_device(self.local)
        self.device = torch.device("cuda", self.local)
        os.environ["OPTSTENCIL_DEVICE"] = str(self.local)   # every Dispersion object of this rank: this rank's GPU
        os.environ["OPTSTENCIL_DEVICES"] = str(self.local)

    def connect(self):
        """Join the process group.  Rank 0 does this AFTER its solo sub-records, so that the other ranks wait in the
        rendezvous on the host instead of spinning in a collective on their GPUs."""
        if self.world > 1:
            self.dist.init_process_group("nccl", device_id=self.device)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.device)

    def max_over_ranks(self, values):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def event_pairs(self, n):
        E = self.torch.cuda.Event
        return [(E(enable_timing=True), E(enable_timing=True)) for _ in range(n)]

    def finish(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def c_oracle_rows(dim, N, coeffs, x_range=None):
    """(S, Amin, Amax, nan) per coefficient row from the plain-C oracle (oracle/oracle_dispersion.c, OpenMP), on the
    reference's tables (dispersion.py:339-356 with Y = Z = 1)."""
    oracle_path()
    import c_oracle
    x = np.linspace(0, np.pi, N)
    cos, s2 = np.cos(x), 0.5 * (1. - np.cos(x))
    co = c_oracle.COracle(dim, [cos] * dim, [s2] * dim, [x] * dim)
    x0, x1 = x_range if x_range is not None else (0, N)
    S, lo, hi, nan = co.objective(np.atleast_2d(coeffs), x0, x1)
    return list(zip(S.tolist(), lo.tolist(), hi.tolist(), nan.tolist()))


def parity_against_c_oracle(N, coeffs, rows, got):
    """got: [4, B] device results; rows: indices re-evaluated by the C oracle on the full N^3 grid."""
    ref = c_oracle_rows(3, N, coeffs[rows])
    rel, exact = 0.0, True
    for r, (S, amin, amax, nn) in zip(rows, ref):
        rel = max(rel, abs(got[0][r] - S) / abs(S))
        amin_c = amin if amin < 0 else 0.0          # the device contract is min / max over sqrtarg U {+0}
        exact = exact and (got[1][r] == amin_c) and (got[2][r] == max(amax, 0.0)) and ((got[3][r] > 0) == (nn > 0))
    return dict(n=len(rows), rows=[int(r) for r in rows], max_rel_S=rel, extrema_bitwise=bool(exact),
                oracle="oracle/oracle_dispersion.c (plain C restatement, pinned to the reference's fixtures), full %d^3 grid" % N,
                tolerance="1e-12 relative on S (north_star); Amin / Amax bit for bit")


def run_gpu_arm(a):
    # Keep stdout to the one JSON line: anything libraries print while the communicator comes up (NCCL's
    # version banner, debug output) is sent to stderr until the line itself is printed.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    env = Env(a)
    torch, dist = env.torch, env.dist
    import optimize_stencil_b200 as sb
    from optimize_stencil_b200 import sharding
    from optimize_stencil_b200.extended_stencil import StencilFree3D, Dispersion3D

    rank, world, local, device = env.rank, env.world, env.local, env.device
    solo = run_solo_extras(a, env) if (rank == 0 and not a.no_extras) else {}
    env.connect()

    N, B = a.grid, a.batch
    stencil = StencilFree3D()
    disp = Dispersion3D(1.0, 1.0, N, stencil)   # the drop-in object: NumPy tables exactly as the reference
    ctx = disp._context()
    params = workload_parameters(B)
    coeffs_host = stencil.coefficients_batch(params)
    coeffs_dev = torch.from_numpy(coeffs_host).to(device)
    # a throughput launch: the shorter asin series (SB200_F_FAST), like Dispersion3D.norm_batch picks for this size
    sharded = sharding.ShardedObjective(sharding.device_evaluator(ctx, fast=True), nx=N, mode="candidates")

    peak = sb.fp64_peak_tflops(local, 0.5)   # measured DFMA roofline denominator, same box, same run

    # ---- device-resident timing: K steps, CUDA events on the launching (torch current) stream -------
    for _ in range(a.warmup):
        out = sharded.evaluate(coeffs_dev)
    env.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count
    ev = env.event_pairs(a.steps)
    lo, hi = sharded.my_candidates(B)
    kev = env.event_pairs(a.steps)
    evaluate_local = sharded.evaluate_local

    def timed_local(block, x0, x1, xs, _k=[0]):
        kev[_k[0]][0].record()
        r = evaluate_local(block, x0, x1, xs)
        kev[_k[0]][1].record()
        _k[0] += 1
        return r

    sharded.evaluate_local = timed_local
    env.barrier()
    t_wall0 = time.perf_counter()
    for s in range(a.steps):
        ev[s][0].record()
        out = sharded.evaluate(coeffs_dev)
        ev[s][1].record()
    env.barrier()
    t_wall = time.perf_counter() - t_wall0
    sharded.evaluate_local = evaluate_local
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launch_count - launches0
    dev_ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    kern_ms = sum(e0.elapsed_time(e1) for e0, e1 in kev)
    dev_ms_max, wall_ms_max = env.max_over_ranks([dev_ms, t_wall * 1e3])
    evals_per_step = float(B) * float(N) ** 3
    value = evals_per_step * a.steps / (dev_ms_max * 1e-3)

    # ---- sanity: the result is finite, gate-clean, and identical on every rank -----------------------
    res = out.cpu().numpy()
    assert res.shape == (4, B) and np.all(np.isfinite(res[0])) and np.all(res[3] == 0), "bench result is not clean"
    chk = float(np.sum(res[0]))
    assert env.max_over_ranks([chk])[0] == chk and -env.max_over_ranks([-chk])[0] == chk, "ranks disagree on the result"

    # ---- end to end through the user-facing call: Dispersion3D.norm_batch on HOST parameters ------------
    # Every rank evaluates its shard of the parameter rows (coefficient fill, pinned H2D, kernel, D2H, gate);
    # N > 1: the norms are then all-gathered so that every rank holds the full answer, like the device path.
    sizes = sharding.block_sizes(B, world)
    bmax = max(sizes)
    pin_in = torch.empty(bmax, dtype=torch.float64).pin_memory() if world > 1 else None
    pin_out = torch.empty(world * bmax, dtype=torch.float64).pin_memory() if world > 1 else None

    parts = dict(norm_batch=0.0, kernel=0.0)   # where the end-to-end step goes (seconds, summed over the timed steps)

    def e2e_step():
        t = time.perf_counter()
        mine = disp.norm_batch(params[lo:hi])
        parts["norm_batch"] += time.perf_counter() - t
        parts["kernel"] += ctx.last_batch_ms * 1e-3
        if world == 1:
            return mine
        pin_in[:hi - lo] = torch.from_numpy(mine)
        g_in = pin_in.to(device, non_blocking=True)
        g_out = torch.empty(world * bmax, dtype=torch.float64, device=device)
        dist.all_gather_into_tensor(g_out, g_in)
        pin_out.copy_(g_out, non_blocking=True)
        torch.cuda.synchronize(device)
        full = pin_out.numpy().reshape(world, bmax)
        return np.concatenate([full[r, :sizes[r]] for r in range(world)])

    for _ in range(max(1, min(a.warmup, 2))):
        e2e_step()
    env.barrier()
    parts.update(norm_batch=0.0, kernel=0.0)
    t0 = time.perf_counter()
    for s in range(a.steps):
        norms = e2e_step()
    t_loop = time.perf_counter() - t0
    env.barrier()
    e2e_s, nb_s, k_s, loop_s = env.max_over_ranks([time.perf_counter() - t0, parts["norm_batch"], parts["kernel"], t_loop])
    e2e_parts = dict(norm_batch_ms=nb_s / a.steps * 1e3, of_which_kernel_ms=k_s / a.steps * 1e3,
                     gather_and_host_ms=(loop_s - nb_s) / a.steps * 1e3,
                     note="max over ranks, per step; the all-gather waits for the slowest rank's kernel")
    e2e_value = evals_per_step * a.steps / e2e_s
    scale = (np.pi / (N - 1)) ** 3
    # same candidates, same kernel; the shard-local tiling may differ from the device path's in the last bits
    assert norms.shape == (B,) and np.allclose(norms, res[0] * scale, rtol=1e-13, atol=0), "norm_batch disagrees"
    h2d = sum(sz * 13 * 8 for sz in sizes) + (sum(sizes) * 8 if world > 1 else 0)
    d2h = sum(4 * sz * 8 for sz in sizes) + (world * world * bmax * 8 if world > 1 else 0)

    # ---- parity inside this run: rows of this batch against the plain-C oracle (outside every timed region) ----
    parity = None
    if rank == 0 and not a.no_parity:
        rows = np.sort(np.random.default_rng(1234).choice(B, size=min(8, B), replace=False))
        parity = parity_against_c_oracle(N, coeffs_host, rows, res)
        assert parity["max_rel_S"] <= 1e-12 and parity["extrema_bitwise"], parity

    extra = dict(run_extras(a, env, sb, peak), **solo) if not a.no_extras else None

    if rank != 0:
        env.finish()
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
                    traffic=traffic, kernel="sb200::objective_kernel_fa<4,true,0,false,0>",
                    flop_per_eval=F_ALG_3D, evals_per_launch=launch_evals, launch_ms=launch_ms,
                    peak_source="DFMA microbenchmark (sb200_fp64_peak: best of several chains x resident-warps shapes) "
                                "measured in this run on this GPU; MEASURED_PEAKS.json has no FP64 entry",
                    peak_nominal=FP64_NOMINAL_TFLOPS, frac_of_nominal=achieved / FP64_NOMINAL_TFLOPS,
                    note="FP64-pipe bound, not HBM or tensor: inputs are O(N) tables + 13 doubles per candidate. "
                         "`achieved` prices every evaluation at the reference formulation's 61 flop (SURVEY.md 8d); "
                         "the kernel regroups sqrtarg and evaluates asin without a square root in the common class, "
                         "so it EXECUTES far fewer FP64 instructions than that (see `executed`) and frac can exceed 1")
    if inst:
        # what the FP64 pipe really does: executed thread-instructions per evaluation (ncu, profiles/) x the
        # evaluation rate of this run, against the issue rate of the pipe (one warp instruction per 2 cycles per
        # SM sub-partition = peak TFLOP/s / 2 flop per DFMA)
        rate = launch_evals / (launch_ms * 1e-3) * inst
        roofline["executed"] = dict(fp64_inst_per_eval=inst, fp64_inst_per_s=rate,
                                    pipe_frac_of_measured=rate / (peak * 1e12 / 2.0),
                                    pipe_frac_of_nominal=rate / (FP64_NOMINAL_TFLOPS * 1e12 / 2.0),
                                    source="profiles/traffic.json (ncu inst_executed by opcode of the bench's own launch)")

    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        v, cores, kind, sample, _ = cpu_reference_steps(N, 1, 0, seed=0)
        cpu = dict(value=v, unit=UNIT, cores=cores, kind=kind, sample=sample)

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=a.steps, warmup=a.warmup,
                ms_per_step=dev_ms_max / a.steps, higher_is_better=True, scaling="strong", vs_baseline=None,
                dtype="f64", data="synthetic", config=config_dict(a), clocks=clocks,
                e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, parts=e2e_parts,
                         api="Dispersion3D.norm_batch(params) on host arrays, all %d rows (every rank its shard: "
                             "Stencil.coefficients_batch -> sb200_objective_batch (ctypes C-ABI, pinned H2D / D2H) -> "
                             "the reference's gate)%s; byte counts are sums over the ranks" %
                             (B, "; then one all-gather of the norms" if world > 1 else "")),
                gpu_launches=int(launches) * world, roofline=roofline, parity=parity, cpu_baseline=cpu,
                wall_ms_per_step=wall_ms_max / a.steps,
                fp64_fraction=dict(of_measured_peak=value / world * F_ALG_3D / (peak * 1e12),
                                   of_nominal_peak=value / world * F_ALG_3D / (FP64_NOMINAL_TFLOPS * 1e12)),
                extra=extra)
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    print(json.dumps(line))
    sys.stdout.flush()
    env.finish()
    return 0


# ---------------------------------------------------------------------------------------------------
# The other named configs of BASELINE.json (1, 2, 3, 5) as sub-records.  Each is bounded to a few seconds.
# ---------------------------------------------------------------------------------------------------
def cli_namespace(**over):
    import types
    base = dict(N=3, scan_dt=True, Ngrid_low=50, Ngrid_high=200, dim=2, div_free=False, symmetric_axes=0,
                symmetric_beta=True, Y=1.0, Z=1.0, dt_multiplier=1.0, deltaxrange=[-1, 0.25],
                deltayrange=[-1, 0.25], deltazrange=[-1, 0.25], betaxyrange=[-1, 1], betaxzrange=[-1, 1],
                betayxrange=[-1, 1], betayzrange=[-1, 1], betazxrange=[-1, 1], betazyrange=[-1, 1],
                dtrange=[0.1, 1], weight="equal", weight_params=[], singlecore=True)
    base.update(over)
    return types.SimpleNamespace(**base)


def timed_steps(env, fn, steps=5, warmup=3):
    """fn() enqueues one step on torch's current stream -> (ms per step, max over ranks; last result)."""
    for _ in range(warmup):
        r = fn()
    env.barrier()
    ev = env.event_pairs(steps)
    for s in range(steps):
        ev[s][0].record()
        r = fn()
        ev[s][1].record()
    env.barrier()
    ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    return env.max_over_ranks([ms])[0] / steps, r


def extra_config1(a, env):
    """configs[0]: optimize_stencil.py with no flags (2-D, StencilSymmetric2D, 81 starts, N = 50 -> 200)."""
    from optimize_stencil_b200.extended_stencil import Optimize
    best = None
    for rep in range(3):                    # the first repetition pays for context creation and module load
        with contextlib.redirect_stdout(io.StringIO()):
            t0 = time.perf_counter()
            opt = Optimize(cli_namespace())
            x, fmin = opt.optimize()
            dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    rec = dict(workload="BASELINE.json configs[0]: optimize_stencil.py defaults (dim 2, StencilSymmetric2D, 81 starts, "
                        "Ngrid 50 -> 200), Optimize(args).optimize() in-process, lock-stepped SLSQP searches",
               wall_s=best, x=[float(v) for v in x], norm=float(fmin), scan=opt.scan_stats,
               launches=int(opt.dispersion.stats["launches"]), evaluated=int(opt.dispersion.stats["evaluated"]),
               reference_wall_s="3.3 s with the reference's 8-process pool, 8.4 s --singlecore (BASELINE.md section 2, "
                                "measured on the survey container, not this box)")
    gpath = os.path.join(ROOT, "tests", "golden", "optimise_runs.npz")
    if os.path.exists(gpath):
        z = np.load(gpath)
        rec["reference_x"] = [float(v) for v in z["default_x"]]
        rec["reference_norm"] = float(z["default_norm"])
        rec["norm_minus_reference"] = float(fmin) - float(z["default_norm"])
        rec["max_abs_coefficient_diff"] = float(np.max(np.abs(np.asarray(x) - z["default_x"])))
        rec["tolerance"] = "SciPy SLSQP ftol = 1e-6 on the objective (north_star)"
    return rec


def extra_config2(a, env):
    """configs[1]: calculate_omega.py --lehe --params 0.96 --dim 3 --Ngrid_high 256."""
    oracle_path()
    import np_oracle as O
    from optimize_stencil_b200.extended_stencil import Dispersion3D, StencilSymmetric3D
    N = 256
    x = [0.96, .125, .125, .125, -0.020732076939092736, 0, 0]
    d = Dispersion3D(1.0, 1.0, N, StencilSymmetric3D())
    d.omega(x)                                                          # warm-up: context, staging buffers
    t0 = time.perf_counter()
    om = d.omega(x)                                                     # a fresh NumPy array: first-touch page faults
    t_omega = time.perf_counter() - t0
    nbytes, secs = d._context().last_export_stats
    t0 = time.perf_counter()
    om2 = d.omega(x, out=om)                                            # the caller's (already touched) array
    t_omega_reuse = time.perf_counter() - t0
    nbytes2, secs2 = d._context().last_export_stats
    assert om2 is om
    torch = env.torch
    dev_buf = torch.empty(N ** 3, dtype=torch.float64, device=env.device)
    pin_buf = torch.empty(N ** 3, dtype=torch.float64).pin_memory()
    pin_buf.copy_(dev_buf); torch.cuda.synchronize(env.device)
    t0 = time.perf_counter()
    pin_buf.copy_(dev_buf, non_blocking=True); torch.cuda.synchronize(env.device)
    pcie = N ** 3 * 8 / (time.perf_counter() - t0) / 1e9
    del dev_buf, pin_buf
    d2 = Dispersion3D(1.0, 1.0, N, StencilSymmetric3D())
    d2.fd_prefetch = False
    d2.norm(x)
    d2._cache.clear()
    t0 = time.perf_counter()
    nrm = d2.norm(x)
    t_norm = time.perf_counter() - t0
    g = O.Grid(3, N)
    ref = O.evaluate(g, O.coefficients("StencilSymmetric3D", x))
    nz = ref["omega"] != 0
    err = float(np.max(np.abs(om - ref["omega"])[nz] / np.abs(ref["omega"][nz])))
    assert err <= 1e-13 and abs(nrm - ref["norm"]) <= 1e-12 * abs(ref["norm"]), (err, nrm, ref["norm"])
    return dict(workload="BASELINE.json configs[1]: calculate_omega.py --lehe --params 0.96 --dim 3 --Ngrid_high 256",
                omega_max_rel_err=err, omega_tolerance=1e-13, norm=float(nrm), norm_oracle=float(ref["norm"]),
                norm_rel_err=abs(nrm - ref["norm"]) / abs(ref["norm"]), norm_tolerance=1e-12,
                norm_reference_digits="48.16521688575889 (SURVEY.md 8c)",
                omega_wall_ms=1e3 * t_omega, omega_export_GBps=nbytes / secs / 1e9,
                omega_into_existing_array_wall_ms=1e3 * t_omega_reuse, omega_into_existing_array_GBps=nbytes2 / secs2 / 1e9,
                pcie_d2h_pinned_GBps=pcie,
                omega_export_note="Dispersion3D.omega(): kernel + chunked double-buffered pinned D2H + host threads copying "
                                  "into the caller's NumPy array (128 MiB); rate = bytes delivered / wall time inside "
                                  "sb200_export.  A FRESH array pays one page fault per 4 KiB on first touch (the host "
                                  "side then limits the rate); omega(x, out=array) fills an existing array at close to the "
                                  "plain pinned D2H rate measured beside it",
                norm_wall_ms=1e3 * t_norm, norm_evals_per_s=N ** 3 / t_norm,
                oracle="oracle/np_oracle.py (NumPy restatement pinned to the reference's fixtures), full 256^3 field")


def extra_config3_scan(a, env, sb):
    """configs[2] (i): the whole 64 x 64 start grid of `--div-free --dim 3 --dtrange 0.6718 1.0` on 128^3 in one
    candidate-sharded launch per rank."""
    torch = env.torch
    from optimize_stencil_b200 import sharding
    from optimize_stencil_b200.extended_stencil import Dispersion3D, StencilIsotropicDivFree3D
    N, n = 128, 64
    st = StencilIsotropicDivFree3D()
    d = Dispersion3D(1.0, 1.0, N, st)
    dts = np.linspace(0.6718, 1.0, n + 2)[1:-1]
    dxs = np.linspace(-1, 0.25, n + 2)[1:-1]
    P = np.array([(t, x) for t in dts for x in dxs])
    C = st.coefficients_batch(P)
    ctx = d._context()
    Cd = torch.from_numpy(C).to(env.device)
    out = {}
    for name, scan in (("plain", False), ("screening", True)):
        sh = sharding.ShardedObjective(sharding.device_evaluator(ctx, scan=scan, fast=True), nx=N, mode="candidates")
        ms, res = timed_steps(env, lambda: sh.evaluate(Cd))
        out[name] = (ms, res.cpu().numpy())
    a0, a1 = out["plain"][1], out["screening"][1]
    # S, Amin, Amax bit for bit; nan is a flag (> 0 iff some omega is NaN): its magnitude depends on the path taken
    same = bool(np.array_equal(a0[1:3], a1[1:3]) and np.array_equal(a0[3] > 0, a1[3] > 0) and
                np.array_equal(np.isnan(a0[0]), np.isnan(a1[0])) and
                np.array_equal(a0[0][~np.isnan(a0[0])], a1[0][~np.isnan(a1[0])]))
    rec = dict(workload="BASELINE.json configs[2] (i): 64 dt x 64 delta_x start grid of --div-free --dim 3 "
                        "--dtrange 0.6718 1.0 (StencilIsotropicDivFree3D) on 128^3, one launch per rank, "
                        "candidate-sharded",
               candidates=len(P), grid=[N] * 3, infeasible_fraction=float(np.mean(np.isnan(a0[0]))),
               ms_plain_kernel=out["plain"][0], ms_screening_kernel=out["screening"][0],
               value=len(P) * float(N) ** 3 / (out["screening"][0] * 1e-3), unit=UNIT,
               variants_identical=same)
    if env.rank == 0:
        rows = np.sort(np.random.default_rng(7).choice(len(P), size=12, replace=False))
        ref = c_oracle_rows(3, N, C[rows])
        ok, rel = True, 0.0
        for r, (S, amin, amax, nn) in zip(rows, ref):
            ok = ok and a1[1][r] == (amin if amin < 0 else 0.0) and a1[2][r] == max(amax, 0.0) and (a1[3][r] > 0) == (nn > 0)
            if nn == 0:
                rel = max(rel, abs(a1[0][r] - S) / abs(S))
        rec["parity"] = dict(n=len(rows), extrema_bitwise=bool(ok), max_rel_S_feasible_rows=rel,
                             oracle="oracle/oracle_dispersion.c")
        assert ok and rel <= 1e-12 and same, rec
        t0 = time.perf_counter()                                   # the user-facing call on host arrays, this rank alone
        res = d.evaluate_batch(P)
        rec["evaluate_batch_wall_ms_one_gpu"] = 1e3 * (time.perf_counter() - t0)
        rec["gated_fraction"] = float(np.mean(res["norm"] == 1e6))
    return rec


def extra_config3_optimize(a, env):
    """configs[2] (ii): the SLSQP scan itself -- 64 dt x 64 delta_x = 4096 searches on 128^3 (16 x 16 with --quick-extras).
    SciPy-only worker processes on the host cores, served by this process on ONE GPU (this rank's)."""
    from optimize_stencil_b200.extended_stencil import Optimize
    n = 16 if a.quick_extras else 64
    with contextlib.redirect_stdout(io.StringIO()):
        t0 = time.perf_counter()
        opt = Optimize(cli_namespace(dim=3, div_free=True, dtrange=[0.6718, 1.0], N=n, Ngrid_low=128, Ngrid_high=128,
                                     singlecore=False))
        x, fmin = opt.optimize()
        dt = time.perf_counter() - t0
    rows = opt.scan_stats.get("rows", opt.dispersion.stats["evaluated"])
    return dict(workload="BASELINE.json configs[2] (ii): optimize_stencil.py --div-free --dim 3 --dtrange 0.6718 1.0 "
                         "--N %d --Ngrid_low 128 --Ngrid_high 128 (%d SLSQP searches on 128^3), one GPU; every search "
                         "bit-identical to its stand-alone run" % (n, n * n),
                wall_s=dt, x=[float(v) for v in x], norm=float(fmin), scan=opt.scan_stats,
                rows_evaluated=int(rows), evals_per_s=rows * 128.0 ** 3 / dt,
                reference_note="the reference's docs call a smaller scan of this kind 'a few minutes even on a very fast "
                               "machine' (command-line-utilities.rst:51-56); round 1 of this repository: 26.8 s")


def extra_config5(a, env, sb):
    """configs[4]: one candidate on 2048^3, k-slab-sharded; all-gather of the 4-double partials + fixed-order combine."""
    torch = env.torch
    from optimize_stencil_b200 import sharding
    from optimize_stencil_b200.extended_stencil import Dispersion3D, StencilFree3D
    N = a.grid5
    st = StencilFree3D()
    d = Dispersion3D(1.0, 1.0, N, st)
    ctx = d._context()
    C = st.coefficients_batch(XC[None, :])
    Cd = torch.from_numpy(C).to(env.device)
    rec = dict(workload="BASELINE.json configs[4]: one candidate (x_c) on a %d^3 k-grid, k-slab-sharded over %d GPU(s): "
                        "rank g owns planes g, g+G, ... ; one NCCL all-gather of 4 doubles per rank + one combine kernel"
                        % (N, env.world), grid=[N] * 3, candidates=1, unit=UNIT)
    results = {}
    for mode in (("slabs", "slabs-contiguous") if env.world > 1 else ("slabs",)):
        sh = sharding.ShardedObjective(sharding.device_evaluator(ctx, fast=True), nx=N, mode=mode,
                                       combine=sharding.device_combiner(ctx))
        ms, res = timed_steps(env, lambda: sh.evaluate(Cd), steps=10, warmup=3)
        results[mode] = res.cpu().numpy()[:, 0]
        rec["ms_" + mode.replace("-", "_")] = ms
    rec["value"] = float(N) ** 3 / (rec["ms_slabs"] * 1e-3)
    rec["fp64_fraction_of_nominal"] = rec["value"] / env.world * F_ALG_3D / (FP64_NOMINAL_TFLOPS * 1e12)
    got = results["slabs"]
    if env.world > 1:
        other = results["slabs-contiguous"]
        assert got[1] == other[1] and got[2] == other[2] and abs(got[0] - other[0]) <= 1e-13 * abs(got[0]), (got, other)
    if env.rank == 0:
        # parity: a sample of x-plane ranges through the same entry point on this rank alone against the C oracle,
        # and the sharded sum against this rank's own whole-grid evaluation
        S1, lo1, hi1, nn1 = ctx.objective_batch(C, fast=True)
        rec["sharded_vs_one_gpu_rel"] = abs(got[0] - S1[0]) / abs(S1[0])
        assert got[1] == lo1[0] and got[2] == hi1[0] and rec["sharded_vs_one_gpu_rel"] <= 1e-13, (got, S1, lo1, hi1)
        rel, exact = 0.0, True
        for xb in (0, N // 3, N - 16):
            S, lo_, hi_, nn = ctx.objective_batch(C, x_begin=xb, x_end=xb + 16)
            (Sr, amin, amax, nnr), = c_oracle_rows(3, N, C, x_range=(xb, xb + 16))
            rel = max(rel, abs(S[0] - Sr) / abs(Sr))
            exact = exact and lo_[0] == (amin if amin < 0 else 0.0) and hi_[0] == max(amax, 0.0)
        rec["parity"] = dict(planes="[0,16), [N/3,N/3+16), [N-16,N)", max_rel_S=rel, extrema_bitwise=bool(exact),
                             oracle="oracle/oracle_dispersion.c on the same plane ranges")
        assert rel <= 1e-12 and exact, rec
        rec["norm"] = float(got[0] * (np.pi / (N - 1)) ** 3)
    return rec


def _guarded(steps, collective=False):
    out = {}
    for name, fn in steps:
        try:
            out[name] = fn()
        except AssertionError:
            raise
        except Exception as exc:             # a sub-record must never cost the headline line ...
            if collective:
                raise                        # ... unless it holds collectives: the ranks must fail together
            out[name] = dict(error=repr(exc))
    return out


def run_solo_extras(a, env):
    """Sub-records that need one GPU and the host cores: rank 0 alone, before it joins the process group."""
    return _guarded([("config1", lambda: extra_config1(a, env)), ("config2", lambda: extra_config2(a, env)),
                     ("config3_optimize", lambda: extra_config3_optimize(a, env))])


def run_extras(a, env, sb, peak):
    """Sub-records every rank takes part in (candidate- and k-slab-sharded)."""
    extra = _guarded([("config3_scan", lambda: extra_config3_scan(a, env, sb)),
                      ("config5_slabs", lambda: extra_config5(a, env, sb))], collective=env.world > 1)
    env.barrier()
    return extra


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--batch", type=int, default=4096, help="candidates per step (default: the named config)")
    ap.add_argument("--grid", type=int, default=512, help="k-grid points per axis (default: the named config)")
    ap.add_argument("--grid5", type=int, default=2048, help="grid of the config-5 sub-record")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the sub-records of the other named configs")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--quick-extras", action="store_true", help="config 3 (ii) with 256 instead of 4096 searches")
    a = ap.parse_args()
    if a.warmup < 3 and a.impl == "b200":
        a.warmup = 3
    return run_reference_arm(a) if a.impl == "reference" else run_gpu_arm(a)


if __name__ == "__main__":
    sys.exit(main())
