"""CPU: the plain-C oracle is pinned against fixtures produced by the unmodified reference."""
import numpy as np
import pytest

import _parity as P


@pytest.mark.parametrize("name", P.golden_names())
def test_oracle_matches_reference_golden(name, oracle):
    meta, gold, inp = P.load_golden(name)
    snaps = P.golden_snapshots(meta, inp, oracle)
    field = "scalar" if meta["nv"] == 1 else "vector"
    tr = oracle.track(snaps, meta["dims"], field=field, jacobian_symmetric=meta["symmetric"])
    # the oracle restates the same arithmetic on the same libm: require bit equality of floats
    P.assert_same_result(P.oracle_result(tr), gold, tol=0.0, what=name)


# counts / digests observed from the reference itself
SURVEY_GOLDENS = {
    "woven_128x128x10": (7359, 422, 66, 0xe7b1a1eb9d4f3c22),
    "woven_10x10x20": (1409, 429, 30, 0x8244fc14b1f829fc),
    "mx2d_11x13x20": (59, 20, 1, 0x99d0d343e22e9760),
    "mx2d_21x21x32": (94, 32, 1, 0x16781bbf4100e385),
    "woven_cli_31x37x32": (4491, 1205, 56, 0xc32490115b6f8871),
    "double_gyre_64x32x50": (879, 100, 2, 0xe9c7ac2d718c30c3),
    "mx3d_21x21x21x10": (38, 10, 1, 0x830858b7b12d52ff),
    "abc_24x24x24x4": (54, 16, 4, 0xd489a8f2c0162f5a),
}


@pytest.mark.parametrize("name", sorted(SURVEY_GOLDENS))
def test_fixture_matches_survey_digest(name):
    npts, nord, ntraj, digest = SURVEY_GOLDENS[name]
    meta, gold, _ = P.load_golden(name)
    pts = gold["points"]
    assert len(pts) == npts and int(pts["ordinal"].sum()) == nord and len(gold["trajectories"]) == ntraj
    assert P.fnv1a64_points(pts, meta["nd"]) == digest


def test_reference_known_answer_counts():
    """Trajectory counts asserted by the reference's own tests."""
    assert len(P.load_golden("woven_cli_31x37x32")[1]["trajectories"]) == 56   # test_critical_point_tracking_woven.cpp:8
    assert len(P.load_golden("double_gyre_64x32x50")[1]["trajectories"]) == 2  # ..._double_gyre.cpp:20-29
    assert len(P.load_golden("woven_10x10x20")[1]["trajectories"]) == 30       # tests/test_critical_point_tracking.py:11-13
    assert len(P.load_golden("mx2d_11x13x20")[1]["trajectories"]) == 1         # tests/test_critical_point_tracking.py:5-9
    assert len(P.load_golden("mx3d_21x21x21x10")[1]["trajectories"]) == 1      # ..._moving_extremum_3d.cpp


def test_moving_extremum_line_equation(oracle):
    """x = x0 + dir * t along the trajectory (reference test_critical_point_tracking_moving_extremum_2d.cpp:23-95)."""
    meta, gold, _ = P.load_golden("mx2d_21x21x32")
    p = gold["points"]
    assert np.abs(p["x"][:, 0] - (10 + 0.1 * p["t"])).max() < 1e-9
    assert np.abs(p["x"][:, 1] - (10 + 0.1 * p["t"])).max() < 1e-9


def test_mesh_table_counts(oracle):
    L = oracle.lib()
    # 3D mesh 1/7/12/6, 4D mesh 1/15/50/60/24; ordinal/interval split of n-simplices
    assert [L.cpo_mesh_ntypes(3, k, 0) for k in range(4)] == [1, 7, 12, 6]
    assert [L.cpo_mesh_ntypes(4, k, 0) for k in range(5)] == [1, 15, 50, 60, 24]
    assert (L.cpo_mesh_ntypes(3, 2, 1), L.cpo_mesh_ntypes(3, 2, 2)) == (2, 10)
    assert (L.cpo_mesh_ntypes(4, 3, 1), L.cpo_mesh_ntypes(4, 3, 2)) == (6, 54)
    assert [L.cpo_mesh_scope_type(3, 2, 1, i) for i in range(2)] == [4, 8]
    assert [L.cpo_mesh_scope_type(4, 3, 1, i) for i in range(6)] == [16, 20, 30, 34, 46, 50]
