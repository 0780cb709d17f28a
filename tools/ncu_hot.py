#!/usr/bin/env python
"""Top stall sites (SASS) per kernel from an ncu report captured with --import-source on.
   python tools/ncu_hot.py prof.ncu-rep [topN] [kernel-substring]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
filt = sys.argv[3] if len(sys.argv) > 3 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kernels, cur = [], None
for row in csv.reader(out.splitlines()):
    if not row:
        continue
    if row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        kernels.append(cur)
    elif row[0] == "Address":
        cur["hdr"] = row
    elif cur is not None and cur["hdr"] is not None:
        cur["rows"].append(row)
for k in kernels:
    if filt not in k["name"]:
        continue
    h = {n: i for i, n in enumerate(k["hdr"])}
    si = h["# Samples"]
    stall_cols = [n for n in k["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(int(r[si] or 0) for r in k["rows"])
    print(f"== {k['name'][:110]}  total samples {tot}")
    agg = {c: sum(int(r[h[c]] or 0) for r in k["rows"]) for c in stall_cols}
    print("   stall mix: " + ", ".join(f"{c[6:]} {100 * v / max(tot, 1):.0f}%" for c, v in sorted(agg.items(), key=lambda x: -x[1])[:7]))
    rows = sorted(enumerate(k["rows"]), key=lambda x: -int(x[1][si] or 0))[:top]
    for idx, r in rows:
        n = int(r[si] or 0)
        why = sorted(((int(r[h[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
        print(f"   {100 * n / max(tot, 1):5.1f}%  #{idx:5d} {r[h['Source']].strip()[:70]:70s} {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}")
