"""ncu --csv log (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum of the GEMM launches of training steps) -> the
per-launch DRAM traffic bench.py reports as roofline.traffic.
usage: python tools/ncu_gemm_traffic.py gpurun_out/gemm_traffic.csv profiles/ncu_r01_gemm_traffic.json --batch 64 --backbone vit-b16"""
import argparse
import collections
import csv
import json

ap = argparse.ArgumentParser()
ap.add_argument('log')
ap.add_argument('out')
ap.add_argument('--batch', type=int, default=64)
ap.add_argument('--backbone', default='vit-b16')
ap.add_argument('--dtype', default='bf16')
a = ap.parse_args()
lines = [l for l in open(a.log) if l.startswith('"')]
per = collections.OrderedDict()
for row in csv.DictReader(lines):
    v = row['Metric Value'].replace(',', '')
    scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'usecond': 1e-6, 'us': 1e-6, 'msecond': 1e-3, 'ms': 1e-3, 'nsecond': 1e-9, 'ns': 1e-9, 'second': 1.0, 's': 1.0}.get(row['Metric Unit'], 1.0)
    per.setdefault(row['ID'], {'name': row['Kernel Name']})[row['Metric Name']] = float(v) * scale
launches = [m for m in per.values() if 'gemm' in m['name']]
n = len(launches)
rd = sum(m['dram__bytes_read.sum'] for m in launches)
wr = sum(m['dram__bytes_write.sum'] for m in launches)
t = sum(m['gpu__time_duration.sum'] for m in launches)
out = dict(kernel='gemm_bf16_sm100_kernel', batch=a.batch, backbone=a.backbone, dtype=a.dtype, launches=n, dram_bytes_read_per_launch=rd / n, dram_bytes_write_per_launch=wr / n,
           dram_bytes_per_launch=(rd + wr) / n, avg_us_under_ncu=t / n * 1e6,
           source='ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gemm_bf16 over the GEMM launches of training steps '
                  f'(python tools/one_step.py --batch {a.batch})')
json.dump(out, open(a.out, 'w'), indent=1)
print(json.dumps(out, indent=1))
