"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total and share of the
window (optionally cut to one learner step: first k_lr_relayout .. following k_lr_adam).
    python scripts/launch_summary.py launches.csv [--learner-step]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
recs = [(r[ki], r[gi], float(r[vi].replace(",", "")) / 1e3) for r in rows[1:] if r[vi].replace(",", "").replace(".", "").isdigit()]
if "--learner-step" in sys.argv:
    a = next(i for i, r in enumerate(recs) if "k_lr_relayout" in r[0])
    b = next(i for i in range(a, len(recs)) if "k_lr_adam" in recs[i][0])
    recs = recs[a:b + 1]
tot = sum(r[2] for r in recs)
agg = collections.OrderedDict()
for name, grid, us in recs:
    short = name.split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    n, t = agg.get(short, (0, 0.0))
    agg[short] = (n + 1, t + us)
print("%-34s %8s %10s %7s" % ("kernel", "launches", "total us", "share"))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-34s %8d %10.1f %6.1f%%" % (k[:34], n, t, 100 * t / tot))
print("%-34s %8d %10.1f" % ("window", len(recs), tot))
