"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
n = 0
for r in rows[hi + 1:]:
    if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
        continue
    n += 1
    if n <= skip:
        continue
    name = re.sub(r"\(.*", "", r[kn])
    name = re.sub(r"^void ", "", name)[:100]
    t = float(r[mv].replace(",", ""))
    agg[name][0] += 1
    agg[name][1] += t
    tot += t
print(f"total {tot / 1e3:.1f} us over {sum(v[0] for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print(f"{v[1] / 1e3:10.1f} us {v[0]:5d}  {100 * v[1] / tot:5.1f}%  {v[1] / v[0] / 1e3:8.1f} us/launch  {k}")
