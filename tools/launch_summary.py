"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total us."""
import csv, sys, collections
path = sys.argv[1]; last = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
rows = rows[1:]
if last: rows = rows[-last:]
agg = collections.OrderedDict()
for r in rows:
    v = float(r[vi].replace(",", "")); v = v / 1000 if r[ui] in ("ns", "nsecond") else v
    k = r[ki][:90]
    c, t = agg.get(k, (0, 0.0)); agg[k] = (c + 1, t + v)
tot = sum(t for _, t in agg.values())
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{t:10.1f} us {c:5d}x {100*t/tot:5.1f}%  {k}")
print(f"{tot:10.1f} us total, {sum(c for c,_ in agg.values())} launches")
