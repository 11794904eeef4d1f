#!/usr/bin/env python
"""Per-step kernel breakdown from an `ncu --metrics gpu__time_duration.sum --csv` launch list of
bench.py (one rSVD step = the launches between two consecutive jacobi_kernel launches)."""
import collections, csv, re, sys
lines = open(sys.argv[1]).readlines()
start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
rows = []
for r in csv.DictReader(lines[start:]):
    if r['Metric Name'] == 'gpu__time_duration.sum':
        name = re.sub(r'\(.*', '', r['Kernel Name']).replace('void <unnamed>::', '').replace('<unnamed>::', '')
        rows.append((name, float(r['Metric Value'].replace(',', ''))))
jac = [i for i, (n, v) in enumerate(rows) if n.startswith('jacobi')]
seg = rows[jac[0] + 1:jac[1] + 1]
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v in seg:
    agg[n][0] += 1; agg[n][1] += v
tot = sum(v for _, v in agg.values())
print(f"one step: {len(seg)} launches, sum of kernel time {tot/1e6:.2f} ms (cold-cache, serialised: compare shares)")
for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1])[:14]:
    print(f"{v/1e6:9.3f} ms {c:4d} {100*v/tot:5.1f}%  {n}")
if len(sys.argv) > 2:
    print([(n[12:24], round(v / 1e3)) for n, v in seg if n.startswith('house')][:30])
