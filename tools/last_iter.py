"""Prints the launches of the last iteration (after the last L2-flush fill) of an ncu launch-list csv."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
data = rows[hi + 1:]
idx = [i for i, r in enumerate(data) if "FillFunctor<unsigned char" in r[ki]]
tot = 0.0
for r in data[idx[-1] + 1:]:
    us = float(r[vi]) / 1000
    tot += us
    print("%9.1f us  %-18s %s" % (us, r[gi], r[ki][:100]))
print("%9.1f us  total" % tot)
