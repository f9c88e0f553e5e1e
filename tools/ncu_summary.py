"""Summarise an Nsight Compute report of the objective kernel into a small text file for profiles/.

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep EVALS_PER_LAUNCH > profiles/rNN_objective_ncu.txt
Reads the report with `ncu -i ... --page raw/source --csv` (works without a GPU).
"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
evals = float(sys.argv[2]) if len(sys.argv) > 2 else None


def page(name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


raw = page("raw")
hdr, units, rows = raw[0], raw[1], raw[2:]
col = {h: i for i, h in enumerate(hdr)}
r = rows[0]


def get(name):
    return r[col[name]] if name in col else "n/a"


print("# Nsight Compute summary:", rep)
print("kernel            :", get("Kernel Name"), " grid", get("Grid Size"), " block", get("Block Size"))
print("captured launches :", len(rows), "(first shown; ncu replays the kernel, times are not bench values)")
wanted = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "CTAs/SM limit (registers)"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM limit (shared memory)"),
    ("sm__warps_active.avg.per_cycle_active", "active warps per SM"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe instructions %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots used %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA(lite) pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "uniform pipe %"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
]
for key, label in wanted:
    if key in col:
        print("%-34s: %s %s" % (label, r[col[key]], units[col[key]]))
print("\n## warp stall reasons (warps per issue-active cycle)")
for h in hdr:
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        v = r[col[h]]
        try:
            if float(v) >= 0.01:
                print("  %-28s %s" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))
        except ValueError:
            pass

src = page("source", ["--print-source", "sass"])
shdr = None
counts = collections.Counter()
stalls = collections.Counter()
for row in src:
    if row and row[0] == "Address":
        if shdr is not None:
            break
        shdr = {h: i for i, h in enumerate(row)}
        continue
    if shdr is None or len(row) < len(shdr):
        continue
    ins = row[shdr["Source"]].split()
    if not ins:
        continue
    op = ins[1] if ins[0].startswith("@") else ins[0]
    op = op.split(".")[0].rstrip(";")
    counts[op] += int(row[shdr["Thread Instructions Executed"]])
    for k in shdr:
        if k.startswith("stall_") and "(" not in k:
            stalls[k] += int(row[shdr[k]])
total = sum(counts.values())
print("\n## executed thread-instructions by opcode (whole launch)")
dp = 0
for op, n in counts.most_common(24):
    line = "  %-10s %16d  %5.1f%%" % (op, n, 100.0 * n / total)
    if evals:
        line += "   %7.2f per evaluation" % (n / evals)
    print(line)
    if op in ("DFMA", "DMUL", "DADD", "DSETP"):
        dp += n
print("  %-10s %16d  %5.1f%%" % ("FP64 total", dp, 100.0 * dp / total) + ("   %7.2f per evaluation (F_alg counts 61 flop)" % (dp / evals) if evals else ""))
if evals:
    print("  %-10s %16d            %7.2f per evaluation" % ("all", total, total / evals))
print("\n## sampled stall reasons (all samples)")
ts = sum(stalls.values())
for k, n in stalls.most_common(8):
    print("  %-24s %5.1f%%" % (k, 100.0 * n / ts))
