"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.DictReader(lines)
agg = collections.OrderedDict()
tot = 0.0
for row in r:
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    us = v / 1000.0 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1000.0)
    name = re.sub(r'\(.*', '', row['Kernel Name'])
    grid = row.get('Grid Size', '')
    key = name if len(sys.argv) < 3 else f'{name} grid={grid}'
    agg.setdefault(key, [0, 0.0]); agg[key][0] += 1; agg[key][1] += us; tot += us
print(f'total {tot:.1f} us over {sum(c for c, _ in agg.values())} launches')
print('kernel,launches,total_us,share')
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'"{k}",{c},{t:.1f},{100 * t / tot:.1f}%')
