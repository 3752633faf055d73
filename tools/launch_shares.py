"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel (dev tool).
usage: python tools/launch_shares.py launches.csv [substring-to-list-individually]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr, start = r, i
        break
ki, gi, vi = hdr.index('Kernel Name'), hdr.index('Grid Size'), hdr.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for r in rows[start + 2:]:
    if len(r) <= vi:
        continue
    name = r[ki].split('(')[0][:70]
    t = float(r[vi].replace(',', '')) / 1000
    agg[name][0] += 1
    agg[name][1] += t
    tot += t
print(f'total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches')
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print('%10.1f us %5.1f%% n=%4d avg=%8.1f us  %s' % (t, 100 * t / tot, n, t / n, k))
if len(sys.argv) > 2:
    for r in rows[start + 2:]:
        if len(r) > vi and sys.argv[2] in r[ki]:
            print(r[ki][:60], r[gi], r[vi])
