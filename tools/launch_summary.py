"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total time and share per kernel."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]
ix = {k: i for i, k in enumerate(h)}
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) < len(h):
        continue
    name = r[ix['Kernel Name']].split('(')[0][-60:]
    v = float(r[ix['Metric Value']].replace(',', ''))
    u = r[ix['Metric Unit']]
    ms = v / 1e6 if u == 'ns' else v / 1e3 if u == 'us' else v if u == 'ms' else v * 1e3
    tot[name][0] += 1
    tot[name][1] += ms
T = sum(v[1] for v in tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:60s} {v[0]:4d} {v[1]:9.2f} ms {100 * v[1] / T:5.1f}%  {1e3 * v[1] / v[0]:8.1f} us/launch")
