"""Generate tests/golden/coords/*.npz: the UNMODIFIED reference run with physical coordinates
(set_coords_bounds / set_coords_rectilinear / set_coords_explicit, regular_tracker.hh:38-40) through
oracle/_ref/ftk_ref_oracle --coords MODE --coords-file F.

    python tests/golden/make_golden_coords.py
Each fixture: the case (meta), the coordinate data as passed, the input snapshots where they are not a generator's,
and the reference's punctured simplices (corner / type / x / t / scalar / cp_type), sorted by element order.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cp_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "coords")


def cases():
    rng = np.random.default_rng(20261018)
    W, H = 24, 20
    yield "woven2d_bounds", 2, 1, [W, H], 5, "woven", None, ("bounds", [-1.5, 2.5, 10.0, 30.0])
    yield "woven2d_rectilinear", 2, 1, [W, H], 5, "woven", None, ("rectilinear", np.concatenate([np.cumsum(rng.random(W) + 0.1), np.cumsum(rng.random(H) + 0.2)]))
    yield "woven2d_explicit2", 2, 1, [W, H], 5, "woven", None, ("explicit", rng.normal(size=(H, W, 2)))
    yield "woven2d_explicit3", 2, 1, [W, H], 5, "woven", None, ("explicit", rng.normal(size=(H, W, 3)))
    yield "gyre2d_vector_bounds", 2, 2, [32, 16], 6, "double_gyre", None, ("bounds", [0.0, 2.0, 0.0, 1.0])
    dims = [8, 9, 10]
    inp = rng.normal(size=(3, dims[2], dims[1], dims[0]))
    yield "rand3d_bounds", 3, 1, dims, 3, None, inp, ("bounds", [0, 1, -2, 2, 5, 6.5])
    yield "rand3d_rectilinear", 3, 1, dims, 3, None, inp, ("rectilinear", np.concatenate([np.cumsum(rng.random(d) + 0.05) for d in dims]))
    yield "rand3d_explicit", 3, 1, dims, 3, None, inp, ("explicit", rng.normal(size=(dims[1], dims[0], 3)))
    vin = rng.normal(size=(3, 7, 6, 7, 3))
    yield "rand3d_vector_rectilinear", 3, 3, [7, 6, 7], 3, None, vin, ("rectilinear", np.concatenate([np.cumsum(rng.random(d) + 0.05) for d in [7, 6, 7]]))


def main():
    O.build()
    os.makedirs(OUT, exist_ok=True)
    for name, nd, nv, dims, T, gen, inp, coords in cases():
        stats, gold = O.run_reference(nd, nv, dims, T, gen=gen, input_array=inp, coords=coords, trace=False)
        p = gold["points"]
        meta = dict(name=name, nd=nd, nv=nv, dims=dims, T=T, gen=gen, coords_mode=coords[0],
                    reference="hguo/ftk@aa4f2cf9 CPU tracker, non-GMP, g++ -O2 -ffp-contract=off -fwrapv")
        arrays = dict(meta=np.frombuffer(json.dumps(meta).encode(), np.uint8), coords=np.ascontiguousarray(coords[1], np.float64).ravel(),
                      corner=p["corner"].astype(np.int32), simplex_type=p["simplex_type"].astype(np.int8), cp_type=p["cp_type"].astype(np.uint8),
                      x=p["x"].copy(), t=p["t"].copy(), scalar=p["scalar"].copy())
        if inp is not None:
            arrays["input"] = inp
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
        print(f"{name}: {len(p)} points")


if __name__ == "__main__":
    main()
