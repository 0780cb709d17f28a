"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): one act step, per kernel."""
import collections, csv, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
rows = []
for row in csv.DictReader(lines):
    try:
        d = float(row["Metric Value"].replace(",", "")) / (1000 if row["Metric Unit"] == "ns" else 1)
        rows.append((row["Kernel Name"], d, row.get("Grid Size")))
    except Exception:
        pass
idx = [i for i, x in enumerate(rows) if "im2col" in x[0]]
step = rows[idx[-2]:idx[-1]]
agg = collections.OrderedDict()
for name, d, grid in step:
    a = agg.setdefault((name[:60], grid), [0, 0.0]); a[0] += 1; a[1] += d
print("kernels", len(step), "sum us %.1f" % sum(d for _, d, _ in step))
for (k, g), (n, t) in agg.items():
    print(f"{n:4d} {t:8.1f} us avg {t/n:7.2f}  grid {g:>14}  {k}")
