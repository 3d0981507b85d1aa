#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): python scripts/ncu_summary.py rep [--stalls] [--md out.md]"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
out = []
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    out.append(f"## {name}  grid={r[hdr.index('launch__grid_size')]} block={r[hdr.index('launch__block_size')]}\n")
    out.append("| metric | value | unit |\n|---|---|---|")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            out.append(f"| {w} | {r[i]} | {units[i]} |")
    if "--stalls" in sys.argv:
        st = [(float(r[i]) if r[i] else 0.0, h) for i, h in enumerate(hdr) if "average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h]
        for v, h in sorted(st, reverse=True)[:8]:
            out.append(f"| {h} | {v:.3f} | warps/issue |")
    out.append("")
txt = "\n".join(out)
print(txt)
if "--md" in sys.argv:
    open(sys.argv[sys.argv.index("--md") + 1], "w").write(txt + "\n")
