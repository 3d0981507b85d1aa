"""Shared parity helpers (test infrastructure).

A *result* is a dict with
  points : structured array sorted by the reference's element order, fields
           corner(4: x,y,z,t) simplex_type ordinal timestep cp_type x(3) t scalar
  trajectories : list of (index array into points, loop flag)
The comparison follows SURVEY.md App. A10: punctured set and integer attributes bit-exact,
trajectories equal as a set of ordered sequences, |dx|,|dt|,|dscalar| <= 1e-9 in grid units.
"""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-9  # north_star: interpolated x/y/(z)/t and scalar within 1e-9 absolute in grid units


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    n = len(z["t"])
    pts = np.zeros(n, dtype=[("corner", np.int32, 4), ("simplex_type", np.int32), ("ordinal", np.int32),
                             ("timestep", np.int32), ("cp_type", np.uint32), ("x", np.float64, 3),
                             ("t", np.float64), ("scalar", np.float64)])
    for f in ("corner", "simplex_type", "ordinal", "timestep", "cp_type", "x", "t", "scalar"):
        pts[f] = z[f]
    off = z["traj_offsets"]
    trajs = [(z["traj_idx"][off[i]:off[i + 1]].astype(np.int64), bool(z["traj_loop"][i])) for i in range(len(off) - 1)]
    inp = z["input"] if "input" in z.files else None
    return meta, {"points": pts, "trajectories": trajs}, inp


def golden_snapshots(meta, inp, oracle):
    """The input snapshots of a golden case, bit-identical to what the reference consumed."""
    if inp is not None:
        return [np.ascontiguousarray(inp[k]) for k in range(meta["T"])]
    return list(oracle.synthetic_series(meta["gen"], meta["dims"], meta["T"], meta["params"]))


def canonical_trajectories(trajs):
    return sorted((tuple(int(i) for i in idx), bool(loop)) for idx, loop in trajs)


def assert_same_result(got, want, check_trajectories=True, tol=TOL, what=""):
    p, g = got["points"], want["points"]
    assert len(p) == len(g), f"{what}: {len(p)} punctured simplices, expected {len(g)}"
    for f in ("corner", "simplex_type", "ordinal", "timestep", "cp_type"):
        a, b = p[f].astype(np.int64), g[f].astype(np.int64)
        assert np.array_equal(a, b), f"{what}: field {f} differs at {np.nonzero((a != b).reshape(len(p), -1).any(1))[0][:5]}"
    for f in ("x", "t", "scalar"):
        a, b = p[f], g[f]
        both_nan = np.isnan(a) & np.isnan(b)
        same_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
        d = np.where(both_nan | same_inf, 0.0, np.abs(a - b))
        assert not np.isnan(d).any() and (d.max() if d.size else 0.0) <= tol, f"{what}: field {f} max|d|={d.max() if d.size else 0}"
    if check_trajectories:
        a, b = canonical_trajectories(got["trajectories"]), canonical_trajectories(want["trajectories"])
        assert len(a) == len(b), f"{what}: {len(a)} trajectories, expected {len(b)}"
        assert a == b, f"{what}: trajectories differ"


def oracle_result(tr):
    return {"points": tr.points(), "trajectories": tr.trajectories()}


def fnv1a64_points(pts, nd):
    """FNV-1a-64 over sorted (corner words.., type) as 8-byte LE words (SURVEY.md App. B digests)."""
    h = 0xcbf29ce484222325
    for c, ty in zip(pts["corner"], pts["simplex_type"]):
        words = [int(c[0]), int(c[1])] + ([int(c[2])] if nd == 3 else []) + [int(c[3]), int(ty)]
        for w in words:
            for byte in int(w).to_bytes(8, "little", signed=True):
                h ^= byte
                h = (h * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h
