"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py launches.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[ix['Metric Name']] != 'gpu__time_duration.sum':
        continue
    name = r[ix['Kernel Name']].split('(')[0][:70]
    v = float(r[ix['Metric Value']].replace(',', '')); u = r[ix['Metric Unit']]
    v = v / 1000 if u == 'ns' else (v * 1000 if u == 'ms' else v)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':72s} {'n':>5s} {'total us':>11s} {'avg us':>9s} {'share':>6s}")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:72s} {a[0]:5d} {a[1]:11.1f} {a[1]/a[0]:9.1f} {100*a[1]/tot:5.1f}%")
