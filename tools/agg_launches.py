"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel name.
    python tools/agg_launches.py file.csv [skip_first_n_launches]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]
kn, mv = h.index('Kernel Name'), h.index('Metric Value')
L = [(r[kn].split('(')[0][:70], float(r[mv].replace(',', ''))) for r in rows[hi + 2:] if len(r) > mv]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
L = L[skip:]
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v in L:
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v for _, v in L)
for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:70s} {c:5d} {v / 1e6:9.3f} ms {100 * v / tot:5.1f}%")
print(f"{'total':70s} {len(L):5d} {tot / 1e6:9.3f} ms")
