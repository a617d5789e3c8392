"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list:
    python scripts/launch_table.py gpurun_out/launches.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
scale = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    a = agg.setdefault(r[ki][:90], [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
total = sum(t for _, t in agg.values())
print("%12s %6s %6s  kernel" % ("total us", "count", "share"))
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print("%12.1f %6d %5.1f%%  %s" % (t, c, 100 * t / total, n))
