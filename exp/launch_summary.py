"""summarise an ncu --metrics gpu__time_duration.sum CSV launch list: per kernel count / total / last"""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    v = v / 1000 if u in ("ns", "nsecond") else (v * 1000 if u in ("ms", "msecond") else v)
    agg.setdefault(r[ki][:70], []).append(v)
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    print(f"{k:70s} n={len(v):3d} sum={sum(v):10.1f} us ({100*sum(v)/tot:5.1f}%) last={v[-1]:10.1f} us")
