"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: launch_agg.py launches.csv"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    try:
        agg[r[ki][:70]].append(float(r[vi].replace(",", "")))
    except Exception:
        pass
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:70s} n={len(v):5d} sum={sum(v)/1e3:10.1f} us ({100*sum(v)/tot:5.1f} %) mean={sum(v)/len(v)/1e3:8.2f} us max={max(v)/1e3:8.2f} us")
