"""Host-only timing of the streaming grow step (ftkb_online_grow, csrc/online.cpp): a woven 2D series is tracked by the parity
oracle (CPU), its punctured simplices are handed to the grow step one timestep at a time.  Usage: grow_step_timing.py [W] [T] [reps]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from oracle import cp_oracle
import ftk_b200
from ftk_b200 import _lib
from ftk_b200.online import OnlineTracer

W = int(sys.argv[1]) if len(sys.argv) > 1 else 768
T = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
cache = f"/tmp/grow_pts_{W}_{T}.npy"
try:
    pts = np.load(cache)
except Exception:
    t0 = time.time()
    snaps = cp_oracle.synthetic_series("woven", [W, W], T)
    o = cp_oracle.track(snaps, [W, W], field="scalar", trace=False)
    op = o.points()
    pts = np.zeros(len(op), _lib.POINT_DTYPE)
    for name in pts.dtype.names:
        if name in op.dtype.names:
            pts[name] = op[name]
    np.save(cache, pts)
    print(f"oracle: {len(pts)} points in {time.time() - t0:.1f} s", file=sys.stderr)
steps = [np.ascontiguousarray(pts[pts["corner"][:, 3] == t]) for t in range(T)]   # the sweep of interval t emits corner time t
n = sum(len(s) for s in steps)
for prepared in (False, True):      # True includes the host-side stand-in for the device preparation (sort + binary searches)
    best = 1e9
    for r in range(reps):
        tr = OnlineTracer([2, 2], [W - 2, W - 2])
        t0 = time.perf_counter()
        for s in steps:
            tr.grow(s, prepared=prepared)
        dt = time.perf_counter() - t0
        best = min(best, dt)
    print(f"{W}x{W}x{T} prepared={prepared}: {n} points, {len(tr.trajectories())} trajectories, grow {best * 1e3:.2f} ms = {best / n * 1e9:.0f} ns/point")
