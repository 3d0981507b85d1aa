#!/usr/bin/env python
"""Top stalled SASS instructions of an .ncu-rep (source page): python scripts/ncu_hot.py rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ia, isrc, ismp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) <= ismp or not r[ismp]: continue
    data.append((int(r[ismp]), r))
tot = sum(d[0] for d in data)
totex = sum(int(d[1][iex] or 0) for d in data)
print("total samples", tot, "instructions executed", totex)
for n, r in sorted(data, key=lambda d: -d[0])[:N]:
    st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:3]
    print(f"{100*n/tot:5.2f}%  ex={r[iex]:>9}  {r[ia][-5:]}  {r[isrc][:60]:60s} {[(h[6:], v) for v, h in st if v]}")
