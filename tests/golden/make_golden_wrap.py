"""Generate tests/golden/wrap/*.npz from the UNMODIFIED reference: SoS vertex ranks beyond 2^31 and 2^32.

The reference's simulation-of-simplicity tie-break ranks a vertex by its index in the (domain x time) lattice, computed in
uint64 and truncated to int (regular_tracker.hh:188-194, lattice.hh:196-207).  At benchmark sizes (8192^2: 67 059 721
corners per time layer) the ranks pass 2^31 in layer t = 32 (at vertex (290, 194)) and 2^32 in layer t = 64 (at vertex
(578, 386)), so the order of tied vertices flips there.  These cases put an integer-aligned moving extremum (exactly
zero gradient components at grid vertices: ties in every determinant) right on those two vertices, start the tracker at
t = 32 / t = 64 (tracker.hh:40 set_current_timestep), and record what the reference's CPU tracker finds.

    python tests/golden/make_golden_wrap.py          (about 10 min of CPU, 10 GB of RAM; run in the build container)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import cp_oracle as O  # noqa: E402
import make_golden as G  # noqa: E402

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "wrap")

# name, dims, start timestep, T, x0 + dir
CASES = [
    ("wrap31_8192x8192_t32", [8192, 8192], 32, 3, [290.0, 194.0, 0.5, 0.5]),      # rank 2^31 is vertex (290, 194) of layer 32
    ("wrap32_8192x8192_t64", [8192, 8192], 64, 3, [578.0, 386.0, 0.5, 0.25]),     # rank 2^32 is vertex (578, 386) of layer 64
]


def main():
    O.build()
    os.makedirs(HERE, exist_ok=True)
    G.HERE = HERE
    for name, dims, t0, T, params in CASES:
        stats, gold = O.run_reference(2, 1, dims, T, gen="moving_extremum", params=params, start_timestep=t0, nthreads=os.cpu_count())
        meta = dict(name=name, nd=2, nv=1, dims=dims, T=T, gen="moving_extremum", params=params, symmetric=None, start_timestep=t0,
                    reference="hguo/ftk@aa4f2cf9 CPU tracker, non-GMP, g++ -O2 -ffp-contract=off -fwrapv")
        G.save(name, meta, gold)
        print(stats)


if __name__ == "__main__":
    main()
