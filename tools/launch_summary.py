"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and launch count per kernel.
    python tools/launch_summary.py launches.csv [top_n] [--step MARKER]
--step MARKER keeps only the launches from the first kernel whose name contains MARKER up to (not including) the second
one, i.e. exactly one step of a trace whose first kernel is MARKER (the stem's `poincare_fwd_kernel`)."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hdr]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
body = [r for r in rows[hdr + 2:] if len(r) > vi]
if "--step" in sys.argv:
    marker = sys.argv[sys.argv.index("--step") + 1]
    hits = [i for i, r in enumerate(body) if marker in r[ki]]
    body = body[hits[0]:hits[1]] if len(hits) >= 2 else body[hits[0]:]
    sys.argv = [a for a in sys.argv if a not in ("--step", marker)]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in body:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki])[:90]
    agg[name][0] += 1
    agg[name][1] += float(r[vi].replace(",", ""))
tot = sum(v[1] for v in agg.values())
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:n]:
    print(f"{v[1] / 1e3:10.1f} us {v[0]:5d} {100 * v[1] / tot:5.1f}% {k}")
print(f"{tot / 1e3:10.1f} us total, {sum(v[0] for v in agg.values())} launches")
