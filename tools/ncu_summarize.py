"""Summarise an ncu --csv log (one row per kernel launch x metric) into a per-kernel table: launches, total/avg time, share, metrics."""
import csv, sys, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.DictReader(lines)
per = collections.OrderedDict()
for row in r:
    key = (row['ID'], row['Kernel Name'])
    per.setdefault(key, {})[row['Metric Name']] = float(row['Metric Value'].replace(',', '')) if row['Metric Value'] not in ('', 'n/a') else 0.0
half = len(per) // 2 if '--second-half' in sys.argv else 0
agg = collections.OrderedDict()
for i, ((_id, name), m) in enumerate(per.items()):
    if i < half:
        continue
    name = name.split('(')[0].split('<')[0]
    a = agg.setdefault(name, dict(n=0, t=0.0, dram=0.0, sm=0.0, occ=0.0, regs=0, rd=0.0, wr=0.0, tens=0.0))
    t = m.get('gpu__time_duration.sum', 0.0)
    a['n'] += 1; a['t'] += t
    a['dram'] += m.get('dram__throughput.avg.pct_of_peak_sustained_elapsed', 0) * t
    a['sm'] += m.get('sm__throughput.avg.pct_of_peak_sustained_elapsed', 0) * t
    a['occ'] += m.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0) * t
    a['tens'] += m.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0) * t
    a['regs'] = max(a['regs'], int(m.get('launch__registers_per_thread', 0)))
    a['rd'] += m.get('dram__bytes_read.sum', 0); a['wr'] += m.get('dram__bytes_write.sum', 0)
tot = sum(a['t'] for a in agg.values())
print(f'total kernel time {tot/1e6:.3f} ms over {sum(a["n"] for a in agg.values())} launches (time-weighted averages; ncu times are cold-cache, serialised)')
print(f'{"kernel":34s} {"n":>4s} {"total ms":>9s} {"share":>6s} {"avg us":>8s} {"dram%":>6s} {"sm%":>6s} {"occ%":>6s} {"tens%":>6s} {"regs":>5s} {"GB moved":>9s}')
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]['t']):
    t = a['t'] or 1
    print(f'{name[:34]:34s} {a["n"]:4d} {a["t"]/1e6:9.3f} {100*a["t"]/tot:5.1f}% {a["t"]/a["n"]/1e3:8.1f} {a["dram"]/t:6.1f} {a["sm"]/t:6.1f} {a["occ"]/t:6.1f} {a["tens"]/t:6.1f} {a["regs"]:5d} {(a["rd"]+a["wr"])/1e9:9.3f}')
