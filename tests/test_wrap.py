"""SoS vertex ranks past 2^31 and 2^32: at benchmark sizes every tie-break runs on ranks that were truncated
from uint64 to int (regular_tracker.hh:188-194).  tests/golden/wrap/*.npz hold what the unmodified reference finds for an
integer-aligned moving extremum (exact zeros in the gradient: ties in every determinant) sitting on the vertex whose rank
is 2^31 (layer 32) / 2^32 (layer 64) of the 8192^2 benchmark domain; tests/golden/make_golden_wrap.py made them."""
import glob
import json
import os

import numpy as np
import pytest

import _parity as P

WRAP_DIR = os.path.join(P.GOLDEN_DIR, "wrap")
NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(WRAP_DIR, "*.npz")))


def _load(name):
    old = P.GOLDEN_DIR
    P.GOLDEN_DIR = WRAP_DIR
    try:
        return P.load_golden(name)
    finally:
        P.GOLDEN_DIR = old


def test_wrap_fixtures_present_and_cross_the_boundary():
    assert len(NAMES) == 2
    for name in NAMES:
        meta, gold, _ = _load(name)
        nx = meta["dims"][0] - 3
        pts = gold["points"]
        assert len(pts) > 0 and len(gold["trajectories"]) >= 1
        # ranks of the corners of the punctured simplices, as the reference computes them (uint64, then truncated)
        rank = (pts["corner"][:, 0].astype(np.int64) - 2) + nx * (pts["corner"][:, 1].astype(np.int64) - 2) + nx * nx * pts["corner"][:, 3].astype(np.int64)
        limit = 2 ** 31 if meta["start_timestep"] == 32 else 2 ** 32
        # the simplices found reach past the limit (their ranks were truncated), and in the 2^31 case their corners straddle it
        assert rank.max() >= limit, f"{name}: no punctured simplex beyond rank {limit}"
        if limit == 2 ** 31:
            assert rank.min() < limit, f"{name}: the punctured simplices do not straddle rank {limit}"


@pytest.mark.skipif(os.environ.get("FTKB_SLOW_TESTS", "") != "1", reason="8 CPU-minutes per case (1.7e9 simplices): FTKB_SLOW_TESTS=1 runs it")
@pytest.mark.parametrize("name", NAMES[:1])
def test_oracle_matches_reference_on_wrapped_ranks(name, oracle):
    """the C restatement (all host threads) against the reference's fixture, bit for bit"""
    meta, gold, _ = _load(name)
    dims, T, prm = meta["dims"], meta["T"], meta["params"]
    snaps = [oracle.gen_moving_extremum(dims, prm[:2], prm[2:], float(k)) for k in range(T)]
    o = oracle.track(snaps, dims, field="scalar", start_timestep=meta["start_timestep"], nthreads=os.cpu_count() or 1)
    P.assert_same_result(P.oracle_result(o), gold, tol=1e-9, what=name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_matches_reference_on_wrapped_ranks(name):
    import ftk_b200
    meta, gold, _ = _load(name)
    dims, T, prm = meta["dims"], meta["T"], meta["params"]
    tr = ftk_b200.make_tracker(dims, field="scalar", start_timestep=meta["start_timestep"])
    for k in range(T):
        tr.push_synthetic_snapshot(0, prm, float(k))       # moving extremum: the device generator is bit-identical (test_device_generators)
        if k:
            tr.advance_timestep()
        if k == T - 1:
            tr.update_timestep()
    tr.finalize()
    got = {"points": tr.get_discrete_critical_points(), "trajectories": tr.get_trajectory_index()}
    P.assert_same_result(got, gold, tol=1e-9, what=name)
    tr.close()
