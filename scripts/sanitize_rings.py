#!/usr/bin/env python
"""Small runs of every shared-memory ring kernel for compute-sanitizer (racecheck / synccheck / memcheck):
    compute-sanitizer --tool racecheck python scripts/sanitize_rings.py
2D scalar (key build kernel, then the fp32-range build kernel through the poison path), 3D scalar (TMA plane ring), 2D / 3D vector,
deferred steps with the test kernel on its own stream; every result is compared with the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ftk_b200  # noqa: E402
from oracle import cp_oracle as O  # noqa: E402
import _parity as P  # noqa: E402

rng = np.random.default_rng(11)


def smooth(dims, T, nv):
    grids = np.meshgrid(*[np.arange(d, dtype=np.float64) for d in reversed(dims)], indexing="ij")
    out = []
    for k in range(T):
        comps = []
        for c in range(max(nv, 1)):
            f = 0.0
            for q, g in enumerate(grids):
                f = f + np.cos((0.7 + 0.17 * c * (q + 1)) * g + 0.3 * q + 0.11 * k + 1.3 * c) * (1.0 + 0.2 * q)
            comps.append(f + 0.05 * rng.standard_normal(size=f.shape))
        out.append(np.stack(comps, axis=-1) if nv > 1 else comps[0])
    return out


cases = [("2d scalar keys", [500, 70], 4, "scalar", None), ("3d scalar", [132, 70, 9], 3, "scalar", None),
         ("2d vector", [140, 40], 3, "vector", None), ("3d vector", [70, 36, 8], 3, "vector", None),
         ("2d scalar f32 (poisoned)", [200, 40], 3, "scalar", "poison")]
for name, dims, T, field, special in cases:
    nv = 1 if field == "scalar" else len(dims)
    snaps = smooth(dims, T, nv)
    if special == "poison":
        snaps[1][5, 7] = np.inf
    o = O.track(snaps, dims, field=field)
    c = ftk_b200.track(snaps, dims, field=field)
    P.assert_same_result({"points": c.get_discrete_critical_points(), "trajectories": c.get_trajectory_index()}, P.oracle_result(o), tol=1e-9, what=name)
    print(name, dims, "ok:", len(o.points()), "punctured simplices,", c.stats()["kernel_launches"], "launches", flush=True)
    c.close()
