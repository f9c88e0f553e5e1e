"""Build experiment variants of the library into build/variants/ (developer tool).

usage: python tools/build_variants.py name:-DMACRO=V,-DMACRO2=V ...   e.g.  deg9:-DSB200_POLY_DEG=9
The variant is selected at run time with SB200_LIB=build/variants/libstencil_b200_<name>.so."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "optimize_stencil_b200", "csrc")
OUT = os.path.join(ROOT, "build", "variants")
os.makedirs(OUT, exist_ok=True)
procs = []
for spec in sys.argv[1:]:
    name, _, flags = spec.partition(":")
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
           "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + [f for f in flags.split(",") if f] + \
          ["-o", os.path.join(OUT, "libstencil_b200_%s.so" % name), "stencil_b200_capi.cu"]
    procs.append((name, subprocess.Popen(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, p in procs:
    out = p.communicate()[0]
    regs = [l for l in out.splitlines() if "registers" in l or "spill" in l.lower() and "0 bytes spill" not in l]
    key = "objective_kernelILi4ELb1ELi0"
    lines = out.splitlines()
    info = [lines[i + 2].strip() for i, l in enumerate(lines) if "Compiling entry" in l and key in l and i + 2 < len(lines)]
    info8 = [lines[i + 2].strip() for i, l in enumerate(lines) if "Compiling entry" in l and "objective_kernelILi8ELb1ELi0" in l]
    print(name, "rc", p.returncode, "| C4:", info, "| C8:", info8)
    if p.returncode:
        print(out[-2000:])
