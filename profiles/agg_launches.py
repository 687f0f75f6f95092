"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel for the LAST step (between the last two SGD axpy launches).
Usage: python profiles/agg_launches.py file.csv [detail-substring]"""
import csv, collections, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h = rows[0]; ki = h.index('Kernel Name'); vi = h.index('Metric Value'); ui = h.index('Metric Unit'); gi = h.index('Grid Size')
data = []
for r in rows[1:]:
    v = float(r[vi].replace(',', ''))
    if r[ui] == 'ns': v /= 1000
    data.append((r[ki], v, r[gi]))
ax = [i for i, d in enumerate(data) if 'OpAxpy' in d[0]]
start, end = (ax[-2] + 1, ax[-1] + 1) if len(ax) > 1 else (0, len(data))
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v, g in data[start:end]:
    k = n.split('(')[0][-58:]
    agg[k][0] += 1; agg[k][1] += v
tot = sum(d[1] for d in data[start:end])
print("step launches", end - start, "total us %.1f" % tot)
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:60s} n={c:4d} {v:9.1f} us {100 * v / tot:5.1f}%")
if len(sys.argv) > 2:
    for n, v, g in data[start:end]:
        if sys.argv[2] in n: print(f"   {v:8.1f} us grid {g}  {n[:90]}")
