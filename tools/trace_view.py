"""Prints the per-batch timeline of a fused-kernel pipeline trace (SE3_FUSED_TRACE build)."""
import sys
from collections import defaultdict
ev = defaultdict(dict)
names = {1: "P_issue", 2: "P_land", 3: "P_ghfree", 4: "P_done", 5: "I1_see", 6: "I1_iss", 7: "A_see", 8: "A_done", 9: "I2_see",
         10: "I2_done", 11: "T_see"}
t0 = None
for ln in open(sys.argv[1]):
    e, i, c = map(int, ln.split())
    t0 = c if t0 is None else min(t0, c)
    ev[i].setdefault(e, c)
lo, hi = int(sys.argv[2]) if len(sys.argv) > 2 else 20, int(sys.argv[3]) if len(sys.argv) > 3 else 44
print("batch " + " ".join("%8s" % names[k] for k in range(1, 11)))
for i in range(lo, hi):
    print("%5d " % i + " ".join("%8s" % (ev[i][k] - t0 if k in ev[i] else "-") for k in range(1, 11)))
last = max(max(v.values()) for v in ev.values())
print("total clocks", last - t0, "batches", max(i for i in ev if 4 in ev[i]) + 1)
