#!/usr/bin/env python
"""Launch list (markdown) from `ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`:
    python scripts/launch_list.py X.csv "title" > profiles/r02_launches_c2.md"""
import collections
import csv
import sys

path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "launch list")
rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
hdr = rows[0]
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= iv or r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[iu], 1e-3)
    t = tot.setdefault(r[ik], [0, 0.0])
    t[0] += 1
    t[1] += v
total = sum(t[1] for t in tot.values())
print(f"# {title}\n")
print("Source: `ncu --metrics gpu__time_duration.sum --clock-control none --csv` under gpurun (B200). Per-launch times under ncu are cold-cache and")
print("serialised: the kernel's SHARE of the step is what to compare with bench.py's CUDA-event numbers (`kernel_ms_per_step`).\n")
print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k[:100]}` | {n} | {us:.1f} | {us / n:.1f} | {100 * us / total:.1f} % |")
