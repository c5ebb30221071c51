"""Kernel shares of a `ncu --metrics gpu__time_duration.sum --csv` launch list.  Usage: python tools/launch_shares.py launches.csv"""
import csv, sys, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
rd = csv.DictReader(lines)
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(r['Metric Value'].replace(',', ''))
    unit = r.get('Metric Unit', 'ns')
    us = v / 1e3 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1e3)
    name = r['Kernel Name'].split('(')[0]
    agg[name][0] += 1
    agg[name][1] += us
tot = sum(v[1] for v in agg.values())
print(f'total_us {tot:.1f} launches {sum(v[0] for v in agg.values())}')
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f'{100 * us / tot:6.2f}% {us:13.1f} us  n={n:5d}  avg={us / n:9.2f} us  {k[:110]}')
