"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares."""
import collections
import csv
import sys

path = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(open(path, errors="replace")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
d = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    name = r[ki].split("(")[0][:80]
    d[name][0] += 1
    d[name][1] += v
tot = sum(v for _, v in d.values())
print(f"total {tot / 1e3:.1f} us over {sum(n for n, _ in d.values())} launches ({tot / 1e3 / steps:.1f} us/step at {steps:g} steps)")
for k, (n, v) in sorted(d.items(), key=lambda x: -x[1][1]):
    print(f"{v / 1e3 / steps:9.1f} us/step {n / steps:6.1f} launches/step {100 * v / tot:5.1f}%  {k}")
