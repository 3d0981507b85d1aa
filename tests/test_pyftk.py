"""The reference's Python front door by its own name: `import pyftk` (python/pyftk.cpp).

GPU tests replay the reference's own known-answer tests (tests/test_critical_point_tracking.py:5-13: 1 and 30
trajectories) through pyftk.trackers / pyftk.extractors and compare every returned record with the oracle."""
import numpy as np
import pytest


def test_pyftk_module_surface():
    import pyftk
    for name in ("track_critical_points_2d_scalar",):
        assert callable(getattr(pyftk.trackers, name))
    for name in ("extract_critical_points_2d_scalar", "extract_critical_points_2d_vector"):
        assert callable(getattr(pyftk.extractors, name))
    for name in ("spiral_woven", "double_gyre_flow", "moving_extremum"):
        assert callable(getattr(pyftk.synthesizers, name))
    assert pyftk.synthesizers.spiral_woven(10, 10, 20).shape == (1, 10, 10, 20)
    assert pyftk.synthesizers.moving_extremum(11, 13, 20, 5, 5, 0.1, 0.2).shape == (1, 11, 13, 20)
    assert pyftk.synthesizers.double_gyre_flow(8, 6, 3).shape == (2, 8, 6, 3)


def _snapshots(data):
    _, DW, DH, DT = data.shape
    flat = np.ascontiguousarray(data, np.float64).reshape(-1)
    return [flat[k * DW * DH:(k + 1) * DW * DH].reshape(DH, DW) for k in range(DT)], [DW, DH]


def _check_traces(result, oracle, data):
    snaps, dims = _snapshots(data)
    o = oracle.track(snaps, dims, field="scalar")
    pts = o.points()
    want = sorted([(float(pts["x"][i][0]), float(pts["x"][i][1]), float(pts["t"][i])) for i in idx] for idx, _ in o.trajectories())
    got = sorted([(p["x"], p["y"], p["t"]) for p in tr["trace"]] for tr in result)
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert len(a) == len(b) and np.abs(np.array(a) - np.array(b)).max() <= 1e-9
    for tr in result:
        assert tr["length"] == len(tr["trace"]) and all(p["type"] in ("min", "max", "saddle", "degenerate", "unknown") for p in tr["trace"])


@pytest.mark.gpu
def test_pyftk_moving_extremum_2d_one_trajectory(oracle):
    """tests/test_critical_point_tracking.py:5-8"""
    import pyftk
    data = pyftk.synthesizers.moving_extremum(11, 13, 20, 5, 5, 0.1, 0.2)
    result = pyftk.trackers.track_critical_points_2d_scalar(data)
    assert len(result) == 1
    _check_traces(result, oracle, data)


@pytest.mark.gpu
def test_pyftk_spiral_woven_thirty_trajectories(oracle):
    """tests/test_critical_point_tracking.py:10-13"""
    import pyftk
    data = pyftk.synthesizers.spiral_woven(10, 10, 20)
    result = pyftk.trackers.track_critical_points_2d_scalar(data)
    assert len(result) == 30
    _check_traces(result, oracle, data)


@pytest.mark.gpu
def test_pyftk_extractors(oracle):
    """single-snapshot extraction (ordinal sweep only), scalar and vector (python/pyftk.cpp:15-90)"""
    import pyftk
    DW, DH = 48, 40
    s = pyftk.synthesizers.spiral_woven(DW, DH, 2).reshape(-1)[:DW * DH].reshape(DW, DH, 1, 1)      # numpy shape (DW, DH, 1, 1): pyftk.cpp:20-22
    got = pyftk.extractors.extract_critical_points_2d_scalar(s)
    o = oracle.track([s.reshape(-1).reshape(DH, DW)], [DW, DH], field="scalar", trace=False)
    want = o.points()
    assert len(got) == len(want) > 0
    a = sorted((p["x"], p["y"], p["scalar"]) for p in got)
    b = sorted((float(q["x"][0]), float(q["x"][1]), float(q["scalar"])) for q in want)
    assert np.abs(np.array(a) - np.array(b)).max() <= 1e-9
    v = pyftk.synthesizers.double_gyre_flow(DW, DH, 2).reshape(-1)[:2 * DW * DH].reshape(2, DW, DH, 1)
    gotv = pyftk.extractors.extract_critical_points_2d_vector(v)
    ov = oracle.track([v.reshape(-1).reshape(DH, DW, 2)], [DW, DH], field="vector", trace=False)
    wantv = ov.points()
    assert len(gotv) == len(wantv) > 0
    a = sorted((p["x"], p["y"]) for p in gotv)
    b = sorted((float(q["x"][0]), float(q["x"][1])) for q in wantv)
    assert np.abs(np.array(a) - np.array(b)).max() <= 1e-9
