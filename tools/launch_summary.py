"""Summarises an ncu launch list (`--metrics gpu__time_duration.sum --csv`): time share per kernel name.  python tools/launch_summary.py file.csv [skip_first_n]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
d = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1 + skip:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    d[r[ki][:90]][0] += 1
    d[r[ki][:90]][1] += v
tot = sum(v[1] for v in d.values())
print(f"# {sys.argv[1]}: {sum(v[0] for v in d.values())} launches, {tot / 1e3:.1f} us (cold-cache, serialised: compare SHARES)")
for k, v in sorted(d.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"{v[1] / 1e3:10.1f} us {100 * v[1] / tot:5.1f}%  n={v[0]:5d}  avg {v[1] / v[0] / 1e3:8.2f} us  {k}")
