import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ftk_b200
for n in (512, 2048, 4096, 8192):
    x0 = (n/2+0.3, n/2-0.3, 0.1, 0.1)
    tr = ftk_b200.make_tracker([n, n], field="scalar")
    prev = 0
    for k in range(4):
        tr.push_synthetic_snapshot(0, list(x0), float(k))
        if k:
            tr.advance_timestep()
            st = tr.stats(); print(n, "step", k, "refined", st["cells_refined"]-prev, "factor", st["scaling_factor"], "res", st["resolution"], "scan ms", st["last_ms_scan"], "derive ms", st["last_ms_derive"]); prev = st["cells_refined"]
    tr.update_timestep()
    st = tr.stats(); print(n, "final", "refined", st["cells_refined"]-prev, "scan ms", st["last_ms_scan"], "points", st["points"])
    tr.close()
