"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys


def main(path, top=30):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        agg.setdefault(r[ki][:90], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print("total %.1f us over %d launches" % (tot, sum(len(v) for v in agg.values())))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:top]:
        print("%-92s n=%4d tot=%10.1f us avg=%8.1f max=%8.1f %5.1f%%" % (k, len(v), sum(v), sum(v) / len(v), max(v), 100 * sum(v) / tot))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
