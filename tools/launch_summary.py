#!/usr/bin/env python
"""tools/launch_summary.py <ncu launch csv> <first launch id> -- per-kernel totals of the launches with ID >= first (one solve / step)."""
import collections
import csv
import re
import sys

path, first = sys.argv[1], int(sys.argv[2])
lines = [ln for ln in open(path) if ln.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
idx = {h: i for i, h in enumerate(hdr)}
rows = []
for r in rd:
    if len(r) < len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
        continue
    rows.append((int(r[idx["ID"]]), r[idx["Kernel Name"]], r[idx["Grid Size"]], float(r[idx["Metric Value"]].replace(",", ""))))
sel = [r for r in rows if r[0] >= first]
print(f"{len(sel)} launches, {sum(r[3] for r in sel) / 1e3:.1f} us in total (cold-cache, serialised: compare shares)")


def short(n):
    n = re.sub(r"^void ", "", n).replace("opf::", "").replace("<unnamed>::", "")
    m = re.match(r"(\w+)", n)
    e = re.search(r"<(.*)>\(", n)
    return m.group(1) + (" " + e.group(1)[:64] if e else "")


agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in sel:
    g = eval(r[2])
    big = "big" if r[3] > 20000 else "small"
    a = agg[(short(r[1]), big)]
    a[0] += 1
    a[1] += r[3]
    a[2] = max(a[2], r[3])
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print(f"{v[1] / 1e3:9.1f} us {100 * v[1] / tot:5.1f}%  n={v[0]:4d}  avg {v[1] / v[0] / 1e3:7.2f}  max {v[2] / 1e3:7.1f} us  {k[1]:5s} {k[0]}")
small = sum(v[1] for k, v in agg.items() if k[1] == "small")
print(f"launches under 20 us: {sum(v[0] for k, v in agg.items() if k[1] == 'small')} = {small / 1e3:.1f} us ({100 * small / tot:.0f}%)")
