"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total time, share."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline='') as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get('Metric Name') == 'gpu__time_duration.sum':
        rows.append((r['Kernel Name'], float(r['Metric Value']), r['Grid Size'], r['Block Size']))
tot = sum(t for _, t, _, _ in rows)
agg = defaultdict(lambda: [0, 0.0])
for name, t, g, b in rows:
    short = re.sub(r'\(.*$', '', name)
    short = re.sub(r'^void ', '', short)
    agg[short][0] += 1
    agg[short][1] += t
print(f'# {path}: {len(rows)} launches, {tot / 1e6:.2f} ms total device time (ncu: cold-cache, serialised)')
print(f'{"share":>7} {"ms":>10} {"launches":>9} {"avg us":>9}  kernel')
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{100 * t / tot:6.2f}% {t / 1e6:10.3f} {c:9d} {t / c / 1e3:9.1f}  {name}')
