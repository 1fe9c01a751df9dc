"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches / mean us / total / share."""
import collections
import csv
import re
import sys


def main(path, top=40):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        if r[ui] in ("ns", "nsecond"):
            v /= 1e3
        elif r[ui] in ("ms", "msecond"):
            v *= 1e3
        name = re.sub(r"\(.*", "", r[ki])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("kernel,launches,mean_us,total_us,share")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%s,%d,%.1f,%.1f,%.3f" % (name, n, t / n, t, t / tot))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
