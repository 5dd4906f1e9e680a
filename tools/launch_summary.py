"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python tools/launch_summary.py FILE [N]"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[kn] == "Kernel Name":
        continue
    name = re.sub(r"\(.*", "", r[kn])
    v = float(r[mv].replace(",", ""))
    ms = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}[r[mu]] * v
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 16]:
    print("| `%s` | %d | %.3f | %.1f %% |" % (k[:70], a[0], a[1], 100 * a[1] / tot))
print("total %.3f ms over %d launches" % (tot, sum(a[0] for a in agg.values())))
