"""Physical coordinates of the interpolated position: set_coords_bounds / set_coords_rectilinear /
set_coords_explicit (regular_tracker.hh:38-40) -> simplex_coordinates (critical_point_tracker_2d_regular.hh:494-526,
critical_point_tracker_3d_regular.hh:343-379).

tests/golden/coords/*.npz were written by the UNMODIFIED reference (tests/golden/make_golden_coords.py).  Punctured set and
types exact; x / t / scalar within 1e-9 absolute (north_star's tolerance, here in the units of the given coordinates).
"""
import glob
import json
import os

import numpy as np
import pytest

from _parity import GOLDEN_DIR, TOL

COORDS_DIR = os.path.join(GOLDEN_DIR, "coords")
NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(COORDS_DIR, "*.npz")))


def load(name):
    z = np.load(os.path.join(COORDS_DIR, name + ".npz"))
    return json.loads(bytes(z["meta"]).decode()), z


def snapshots(meta, z, oracle):
    if "input" in z.files:
        return [np.ascontiguousarray(z["input"][k]) for k in range(meta["T"])]
    return list(oracle.synthetic_series(meta["gen"], meta["dims"], meta["T"], None))


def check(points, z, what):
    assert len(points) == len(z["t"]), what
    assert np.array_equal(points["corner"], z["corner"]) and np.array_equal(points["simplex_type"], z["simplex_type"]), what
    assert np.array_equal(points["cp_type"].astype(np.int64), z["cp_type"].astype(np.int64)), what
    for f in ("x", "t", "scalar"):
        d = np.abs(points[f] - z[f])
        assert d.size == 0 or d.max() <= TOL, f"{what}: {f} max|d| = {d.max()}"


def test_fixtures_present():
    assert len(NAMES) >= 9
    modes = {load(n)[0]["coords_mode"] for n in NAMES}
    assert modes == {"bounds", "rectilinear", "explicit"}


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference(name, oracle):
    meta, z = load(name)
    field = "scalar" if meta["nv"] == 1 else "vector"
    tr = oracle.track(snapshots(meta, z, oracle), meta["dims"], field=field, trace=False, coords=(meta["coords_mode"], z["coords"]))
    check(tr.points(), z, name)
    # the coordinates really differ from grid units (the fixture is not vacuous)
    plain = oracle.track(snapshots(meta, z, oracle), meta["dims"], field=field, trace=False).points()
    assert np.abs(plain["x"] - z["x"]).max() > 1e-3


def test_set_coords_argument_checks():
    from ftk_b200 import _lib
    assert _lib.lib().ftkb_set_coords(None, 1, None, 0) == 1          # FTKB_ERR_INVALID


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_matches_reference(name, oracle):
    from ftk_b200 import tracker as T
    meta, z = load(name)
    field = "scalar" if meta["nv"] == 1 else "vector"
    tr = T.track(snapshots(meta, z, oracle), meta["dims"], field=field, trace=False, coords=(meta["coords_mode"], z["coords"]))
    check(tr.get_discrete_critical_points(), z, name)
    tr.close()


@pytest.mark.gpu
def test_gpu_set_coords_rejects_wrong_sizes():
    from ftk_b200 import _lib, tracker as T
    tr = T.make_tracker([16, 12], field="scalar")
    bad = np.zeros(5)
    for mode in (_lib.COORDS_BOUNDS, _lib.COORDS_RECTILINEAR, _lib.COORDS_EXPLICIT):
        assert _lib.lib().ftkb_set_coords(tr._h, mode, bad.ctypes.data, bad.size) == 1
    assert _lib.lib().ftkb_set_coords(tr._h, 7, bad.ctypes.data, bad.size) == 1
    tr.close()
