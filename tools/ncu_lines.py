"""Aggregates the warp-stall samples of an ncu report by CUDA source line (needs --import-source on, -lineinfo).
usage: python tools/ncu_lines.py report.ncu-rep [top N]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = defaultdict(lambda: [0, defaultdict(int), ""])
fname, h = "", None
first_kernel_done = False
for r in rows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        if first_kernel_done:
            break
        continue
    if r[0] == "File Name":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        h = r
        isamp = h.index("# Samples")
        stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        continue
    if h is None or len(r) <= isamp:
        continue
    if r[0]:
        cur = (fname, int(r[0]))
        agg[cur][2] = r[1].strip()
        first_kernel_done = True
    try:
        n = int(r[isamp] or 0)
    except ValueError:
        continue
    a = agg[cur]
    a[0] += n
    for c in stall:
        v = int(r[c] or 0)
        if v:
            a[1][h[c][6:]] += v
tot = sum(a[0] for a in agg.values())
print("total samples", tot)
for key in sorted(sorted(agg, key=lambda k: -agg[k][0])[:top]):
    a = agg[key]
    st = sorted(a[1].items(), key=lambda kv: -kv[1])[:3]
    print("%-16s %4d %6d  %-90s %s" % (key[0], key[1], a[0], a[2][:90], " ".join("%s:%d" % kv for kv in st)))
