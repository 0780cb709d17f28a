#!/usr/bin/env python
"""Stall samples of one kernel grouped by code region (consecutive SASS index ranges split at given marker mnemonics):
   python tools/ncu_region.py prof.ncu-rep kernel-substring [bucket]     # bucket = instructions per bucket (default 150)"""
import csv, subprocess, sys
rep, filt = sys.argv[1], sys.argv[2]
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 150
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
cur, hdr, rows = None, None, []
for row in csv.reader(out.splitlines()):
    if not row:
        continue
    if row[0] == "Kernel Name":
        cur = row[1]; hdr = None
    elif row[0] == "Address" and filt in (cur or ""):
        hdr = row; rows = []
    elif hdr is not None and filt in (cur or ""):
        rows.append(row)
h = {n: i for i, n in enumerate(hdr)}
si = h["# Samples"]
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(int(r[si] or 0) for r in rows)
print(f"{len(rows)} instructions, {tot} samples")
for b0 in range(0, len(rows), bucket):
    rs = rows[b0:b0 + bucket]
    n = sum(int(r[si] or 0) for r in rs)
    if n < tot * 0.01:
        continue
    agg = sorted(((sum(int(r[h[c]] or 0) for r in rs), c[6:]) for c in stall_cols), reverse=True)[:4]
    ops = {}
    for r in rs:
        op = r[h["Source"]].strip().split()[0].lstrip("@!P0123456789 ") if r[h["Source"]].strip() else ""
        src = r[h["Source"]].strip()
        op = [t for t in src.split() if not t.startswith("@")][0] if src else ""
        ops[op] = ops.get(op, 0) + 1
    top_ops = ", ".join(f"{k}x{v}" for k, v in sorted(ops.items(), key=lambda x: -x[1])[:5])
    print(f"  [{b0:5d},{b0 + len(rs):5d}) {100 * n / tot:5.1f}%  " + ", ".join(f"{c}:{v}" for v, c in agg) + f"   | {top_ops}")
