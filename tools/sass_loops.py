#!/usr/bin/env python
"""Opcode census of the loops of one kernel in a built library (cuobjdump -sass), no GPU needed.

    python tools/sass_loops.py [lib.so] [--kernel 'objective_kernel_fa<4, true, 0, false>'] [--dump N]

Every backward branch closes a loop [target, branch]; for each loop the script prints the number of FP64-pipe
instructions (DFMA/DMUL/DADD/DSETP/...), of all others, and the issue-slot estimate 2*FP64 + other (B200: a warp-wide
FP64 instruction holds the dispatch port for two cycles, profiles/r01_ubench_issue.log).  The straight-line MID path
of the objective kernel is the innermost loop with the most DFMAs; conditional side paths inside a loop are counted
as if always executed, so treat the figures of the outer loops as upper bounds.
"""
import argparse
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX", "F2F.F64", "I2F.F64", "F2I.F64", "MUFU.RCP64H", "MUFU.RSQ64H")


def sass_of(lib, kernel):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    blocks = re.split(r"\n\s*Function : ", txt)[1:]
    out = {}
    for b in blocks:
        name = b.split("\n", 1)[0].strip()
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        out[dem] = b
    want = [k for k in out if kernel in k]
    if not want:
        raise SystemExit("no kernel matching %r; have:\n  %s" % (kernel, "\n  ".join(sorted(out))))
    return want[0], out[want[0]]


def parse(body):
    ins = []
    for line in body.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if not m:
            continue
        addr, text = int(m.group(1), 16), m.group(2).strip()
        pred = None
        pm = re.match(r"(@!?U?P\d+)\s+(.*)", text)
        if pm:
            pred, text = pm.group(1), pm.group(2)
        op = text.split()[0]
        ins.append((addr, pred, op, text))
    return ins


def is_fp64(op):
    return any(op.startswith(f) for f in FP64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("lib", nargs="?", default=os.path.join(ROOT, "optimize_stencil_b200", "csrc", "libstencil_b200.so"))
    ap.add_argument("--kernel", default="objective_kernel_fa<4, true, 0, false>")
    ap.add_argument("--dump", type=int, default=-1, help="print the instructions of loop number N")
    ap.add_argument("--min-fp64", type=int, default=8)
    a = ap.parse_args()
    name, body = sass_of(a.lib, a.kernel)
    ins = parse(body)
    index = {addr: i for i, (addr, _, _, _) in enumerate(ins)}
    print("kernel:", name, "-", len(ins), "instructions,", sum(is_fp64(o) for _, _, o, _ in ins), "FP64")
    loops = []
    for i, (addr, pred, op, text) in enumerate(ins):
        if op.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", text)
            if m and int(m.group(1), 16) <= addr and int(m.group(1), 16) in index:
                loops.append((index[int(m.group(1), 16)], i))
    loops.sort(key=lambda l: (l[1] - l[0]))
    for n, (lo, hi) in enumerate(loops):
        seg = ins[lo:hi + 1]
        f = sum(is_fp64(o) for _, _, o, _ in seg)
        if f < a.min_fp64:
            continue
        census = collections.Counter(o.split(".")[0] for _, _, o, _ in seg)
        other = len(seg) - f
        inner = [m for m, (l2, h2) in enumerate(loops) if l2 >= lo and h2 <= hi and (l2, h2) != (lo, hi)]
        print("loop %2d  [%05x..%05x]  %4d instr: %3d FP64 + %3d other -> %4d issue slots   inner loops: %s"
              % (n, seg[0][0], seg[-1][0], len(seg), f, other, 2 * f + other, inner or "-"))
        print("         " + "  ".join("%s %d" % kv for kv in census.most_common(14)))
        if n == a.dump:
            for addr, pred, op, text in seg:
                print("    %05x  %-6s %s" % (addr, pred or "", text))


if __name__ == "__main__":
    main()
