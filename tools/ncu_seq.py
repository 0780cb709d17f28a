"""Per-launch durations (us) of the last N launches of an ncu launch list, in order, with short kernel names."""
import csv, sys
path, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = []
for row in csv.DictReader(l for l in open(path) if not l.startswith("==")):
    try:
        d = float(row["Metric Value"].replace(",", "")) / (1000 if row["Metric Unit"] == "ns" else 1)
        rows.append((row["Kernel Name"][:44], d))
    except Exception:
        pass
for k, d in rows[-n:]:
    print(f"{d:8.2f}  {k}")
