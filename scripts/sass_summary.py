#!/usr/bin/env python
"""Per-kernel SASS opcode summary of the in-tree library (cuobjdump -sass, no GPU needed):
    python scripts/sass_summary.py [ftk_b200/libftkb200.so] > profiles/r02_sass_summary.md
Counts the opcodes that identify the staging (UBLKCP = cp.async.bulk, UTMALDG = TMA tensor load, SYNCS = mbarrier),
the arithmetic the scans live on (DADD, FMNMX / FMNMX3, F2F conversions) and local-memory spills (LDL / STL)."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "ftk_b200/libftkb200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kernels, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        kernels[cur][m.group(1)] += 1
demangled = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
cols = ["UBLKCP", "UTMALDG", "SYNCS", "LDS", "LDG", "DADD", "DSETP", "FMNMX", "FMNMX3", "F2F", "SHFL", "ATOMG", "LDL", "STL"]
print("# SASS opcode summary per kernel (sm_100a, `cuobjdump -sass " + lib + "`)\n")
print("UBLKCP = `cp.async.bulk` (bulk async copy into shared memory), UTMALDG = `cp.async.bulk.tensor` (TMA tile load), SYNCS = mbarrier operations,")
print("F2F = fp64->fp32 conversions (XU pipe), LDL/STL = local-memory traffic (spills, stack).  Static instruction counts.\n")
print("| kernel | instructions | " + " | ".join(cols) + " |")
print("|---|---|" + "---|" * len(cols))
for (name, c), dm in sorted(zip(kernels.items(), demangled), key=lambda t: t[1]):
    short = re.sub(r"\(.*", "", dm).replace("ftkb::", "").replace("void ", "")
    if "cub::" in dm or "thrust::" in dm:
        continue
    print(f"| `{short}` | {sum(c.values())} | " + " | ".join(str(c.get(k, 0)) for k in cols) + " |")
