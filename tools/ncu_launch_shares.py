"""ncu --csv launch list (gpu__time_duration.sum per launch) -> per-kernel launches / total time / share table.
usage: python tools/ncu_launch_shares.py <launches.csv> "<header comment>" > profiles/..._shares.txt"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row['Metric Name'] != 'gpu__time_duration.sum':
        continue
    scale = {'usecond': 1.0, 'us': 1.0, 'msecond': 1e3, 'ms': 1e3, 'nsecond': 1e-3, 'ns': 1e-3, 'second': 1e6, 's': 1e6}[row['Metric Unit']]
    name = row['Kernel Name'].split('(')[0].split('<')[0].replace('void ', '')
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(row['Metric Value'].replace(',', '')) * scale
tot = sum(a[1] for a in agg.values())
for c in sys.argv[2:]:
    print('# ' + c)
print(f'# total {tot / 1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches')
print('kernel | launches | total us | share')
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{name} | {n} | {t:.0f} | {100 * t / tot:.1f}%')
