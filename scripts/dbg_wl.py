import sys, numpy as np
sys.path.insert(0, ".")
import ftk_b200
for dims in ([64, 48], [200, 100], [1024, 1024]):
    x0 = [dims[0] / 2 + 0.3, dims[1] / 2 - 0.3]; d = [0.1, 0.1]
    tr = ftk_b200.make_tracker(dims, field="scalar")
    for k in range(2):
        tr.push_synthetic_snapshot(0, x0 + d, float(k))
        if k:
            tr.advance_timestep()
    wl = tr.get_last_worklist()
    nc0 = dims[0] - 3
    xs, ys = wl % nc0 + 2, wl // nc0 + 2
    print(dims, "worklist", len(wl), "x", np.unique(xs)[:40], "y", np.unique(ys)[:40], flush=True)
    print("   stats", {k: v for k, v in tr.stats().items() if k in ("cells_refined", "scaling_factor", "resolution", "sweeps_repeated")})
    tr.close()
