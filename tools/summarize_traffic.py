"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list by
kernel: time, launches, DRAM bytes read / written.  Kernels whose name matches argv[2] (regex) are skipped (e.g. the Gram
build that precedes the factorisation)."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
skip = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
with open(path, newline='') as f:
    lines = [l for l in f if l.startswith('"')]
per = defaultdict(dict)  # launch id -> metric -> (value, unit)
names = {}
for r in csv.DictReader(lines):
    per[r['ID']][r['Metric Name']] = (float(r['Metric Value'].replace(',', '')), r['Metric Unit'])
    names[r['ID']] = r['Kernel Name']
unit = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'nsecond': 1e-6,
        'usecond': 1e-3, 'msecond': 1.0, 'second': 1e3}
agg = defaultdict(lambda: [0.0, 0, 0.0, 0.0])
for i, m in per.items():
    name = re.sub(r'\(.*$', '', names[i])
    name = re.sub(r'^void (lgp::)?', '', name)
    if skip and skip.search(name):
        continue
    t = m['gpu__time_duration.sum'][0] * unit[m['gpu__time_duration.sum'][1]]
    rd = m['dram__bytes_read.sum'][0] * unit[m['dram__bytes_read.sum'][1]]
    wr = m['dram__bytes_write.sum'][0] * unit[m['dram__bytes_write.sum'][1]]
    a = agg[name]
    a[0] += t; a[1] += 1; a[2] += rd; a[3] += wr
T = sum(a[0] for a in agg.values()); R = sum(a[2] for a in agg.values()); W = sum(a[3] for a in agg.values())
print(f'# total: {T:.2f} ms, dram read {R / 1e9:.2f} GB, dram write {W / 1e9:.2f} GB, read + write {(R + W):.6g} bytes')
print(f'{"ms":>10} {"launches":>9} {"read GB":>9} {"write GB":>9}  kernel')
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f'{a[0]:10.3f} {a[1]:9d} {a[2] / 1e9:9.3f} {a[3] / 1e9:9.3f}  {name}')
