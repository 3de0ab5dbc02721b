#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
tot = collections.defaultdict(float); cnt = collections.Counter()
for row in rows:
    name = re.sub(r'\(.*', '', row['Kernel Name'])
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    v = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v
    tot[name] += v; cnt[name] += 1
T = sum(tot.values())
print("%d launches, %.1f ms total" % (len(rows), T / 1e3))
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print("%-62s n=%5d total=%10.1f us avg=%9.1f us share=%5.1f%%" % (k[:62], cnt[k], v, v / cnt[k], 100 * v / T))
