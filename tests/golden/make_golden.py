"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/ftk_ref_oracle).

Run in the build container (needs /root/reference for oracle/build_ref.sh):
    python tests/golden/make_golden.py
Every fixture holds the reference's discrete critical points (sorted by the reference's element
order), its trajectories (CSR over the point array, canonical order) and the case description.
Cases with explicit inputs (random / degenerate fields) also store the input bits.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cp_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name, nd, nv, dims, T, generator, params   (SURVEY.md App. B goldens + reference test configs)
GENERATED = [
    ("woven_128x128x10", 2, 1, [128, 128], 10, "woven", None),           # BASELINE configs[0] (C1)
    ("woven_10x10x20", 2, 1, [10, 10], 20, "woven", None),               # tests/test_critical_point_tracking.py
    ("mx2d_11x13x20", 2, 1, [11, 13], 20, "moving_extremum", [5, 5, 0.1, 0.2]),
    ("mx2d_21x21x32", 2, 1, [21, 21], 32, "moving_extremum", None),
    ("woven_cli_31x37x32", 2, 1, [31, 37], 32, "woven_cli", None),       # tests/test_critical_point_tracking_woven.cpp
    ("merger_32x32x100", 2, 1, [32, 32], 100, "merger", None),
    ("double_gyre_64x32x50", 2, 2, [64, 32], 50, "double_gyre", None),
    ("mx3d_21x21x21x10", 3, 1, [21, 21, 21], 10, "moving_extremum", None),
    ("abc_24x24x24x4", 3, 3, [24, 24, 24], 4, "abc", None),
]


def explicit_cases():
    rng = np.random.default_rng(20261017)
    cases = []
    cases.append(("rand2d_scalar_int", 2, 1, [13, 11], 5, [rng.integers(-2, 3, size=(11, 13)).astype(np.float64) for _ in range(5)], None))
    cases.append(("rand2d_vector_int", 2, 2, [12, 11], 4, [rng.integers(-3, 4, size=(11, 12, 2)).astype(np.float64) for _ in range(4)], None))
    cases.append(("rand2d_vector_normal_sym", 2, 2, [12, 10], 4, [rng.standard_normal(size=(10, 12, 2)) for _ in range(4)], True))
    a = [rng.standard_normal(size=(9, 10, 2)) for _ in range(3)]
    a[1][3, 4, 0] = np.nan
    a[2][5, 5, 1] = np.inf
    a[0][2, 2, :] = 0
    cases.append(("nan_inf_2d_vector", 2, 2, [10, 9], 3, a, None))
    cases.append(("huge_2d_vector", 2, 2, [10, 9], 3, [rng.standard_normal(size=(9, 10, 2)) * 1e13 for _ in range(3)], None))
    cases.append(("rand3d_scalar_int", 3, 1, [8, 7, 6], 3, [rng.integers(-2, 3, size=(6, 7, 8)).astype(np.float64) for _ in range(3)], None))
    cases.append(("rand3d_scalar_normal", 3, 1, [8, 7, 7], 3, [rng.standard_normal(size=(7, 7, 8)) for _ in range(3)], None))
    cases.append(("rand3d_vector_int", 3, 3, [7, 7, 6], 3, [rng.integers(-2, 3, size=(6, 7, 7, 3)).astype(np.float64) for _ in range(3)], None))
    cases.append(("rand3d_vector_normal_sym", 3, 3, [7, 6, 7], 3, [rng.standard_normal(size=(7, 6, 7, 3)) for _ in range(3)], True))
    cases.append(("huge_3d_vector", 3, 3, [6, 6, 6], 2, [rng.standard_normal(size=(6, 6, 6, 3)) * 1e9 for _ in range(2)], None))
    # smooth 3D scalar field with minima, maxima and saddles (exercises the trigonometric eigen-solver)
    D = 14
    z, y, x = np.meshgrid(np.arange(D), np.arange(D), np.arange(D), indexing="ij")
    a = [np.cos(0.9 * x + 0.13 * k) * np.cos(0.8 * y - 0.07 * k) * np.cos(0.85 * z + 0.05 * k) + 0.01 * x for k in range(4)]
    cases.append(("cos3d_14x14x14x4", 3, 1, [D, D, D], 4, a, None))
    return cases


def save(name, meta, gold, input_array=None):
    pts = gold["points"]
    trajs = O.canonical_trajectories(gold["trajectories"])
    off = np.zeros(len(trajs) + 1, np.int64)
    for i, (idx, _) in enumerate(trajs):
        off[i + 1] = off[i] + len(idx)
    idx = np.concatenate([np.asarray(t[0], np.int64) for t in trajs]) if trajs else np.zeros(0, np.int64)
    loop = np.asarray([t[1] for t in trajs], np.uint8)
    arrays = dict(
        meta=np.frombuffer(json.dumps(meta).encode(), np.uint8),
        corner=pts["corner"].astype(np.int32), simplex_type=pts["simplex_type"].astype(np.int8),
        ordinal=pts["ordinal"].astype(np.int8), timestep=pts["timestep"].astype(np.int32),
        cp_type=pts["cp_type"].astype(np.uint8), x=pts["x"].copy(), t=pts["t"].copy(), scalar=pts["scalar"].copy(),
        traj_offsets=off, traj_idx=idx.astype(np.int32), traj_loop=loop)
    if input_array is not None:
        arrays["input"] = input_array
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    print(f"{name}: {len(pts)} points, {len(trajs)} trajectories")


def main():
    O.build()
    for name, nd, nv, dims, T, gen, params in GENERATED:
        stats, gold = O.run_reference(nd, nv, dims, T, gen=gen, params=params)
        meta = dict(name=name, nd=nd, nv=nv, dims=dims, T=T, gen=gen, params=params, symmetric=None,
                    reference="hguo/ftk@aa4f2cf9 CPU tracker, non-GMP, g++ -O2 -ffp-contract=off -fwrapv")
        save(name, meta, gold)
    for name, nd, nv, dims, T, arr, sym in explicit_cases():
        inp = np.stack(arr)
        stats, gold = O.run_reference(nd, nv, dims, T, input_array=inp, symmetric=sym)
        meta = dict(name=name, nd=nd, nv=nv, dims=dims, T=T, gen=None, params=None, symmetric=sym,
                    reference="hguo/ftk@aa4f2cf9 CPU tracker, non-GMP, g++ -O2 -ffp-contract=off -fwrapv")
        save(name, meta, gold, inp)


if __name__ == "__main__":
    main()
