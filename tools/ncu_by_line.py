"""Join an ncu source-page export (SASS rows with executed-instruction counts and stall samples) with nvdisasm's
line info of the same cubin: executed thread instructions and samples per source line.

    ncu -i X.ncu-rep --page source --csv > sass.csv
    cuobjdump -xelf all libstencil_b200.so; nvdisasm --print-line-info *.cubin > all.dis
    python tools/ncu_by_line.py sass.csv all.dis '<mangled kernel name fragment>' [top]
"""
import csv
import re
import sys
from collections import defaultdict

sass_csv, dis, frag = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
col = {n: i for i, n in enumerate(hdr)}
data = rows[2:]
base = int(data[0][col["Address"]], 16)
stat = {}
for r in data:
    off = int(r[col["Address"]], 16) - base
    stat[off] = (r[col["Source"]].strip(), int(r[col["Thread Instructions Executed"]]), int(r[col["# Samples"]]),
                 int(r[col["Instructions Executed"]]))
line_of = {}
cur = None
inside = False
for ln in open(dis):
    if ln.startswith(".text."):
        inside = frag in ln
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
by_line = defaultdict(lambda: [0, 0, 0, 0])
tot_t = tot_s = 0
for off, (src, thr, smp, wi) in stat.items():
    key = line_of.get(off)
    b = by_line[key]
    b[0] += thr; b[1] += smp; b[2] += wi
    if re.match(r"(@!?U?P\d+\s+)?D(FMA|MUL|ADD|SETP)", src):
        b[3] += thr
    tot_t += thr; tot_s += smp
print("total thread instructions %.4g, samples %d" % (tot_t, tot_s))
print("%-34s %8s %8s %8s %6s" % ("source line", "thr-ins%", "samples%", "fp64%", "lanes"))
for key, (thr, smp, wi, f64) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-34s %8.2f %8.2f %8.2f %6.1f" % ("%s:%d" % key if key else "?", 100 * thr / tot_t, 100 * smp / tot_s, 100 * f64 / tot_t, thr / max(wi, 1)))
