"""Several devices behind one tracker (ftkb_group, include/ftkb200.h): the result equals the one-device tracker's bit for
bit -- punctured simplices, trajectories, order.  Device lists may repeat an id, so the chunking / halo / inheritance /
merge logic is exercised on a one-GPU box too; with two or more GPUs visible the same cases run across them."""
import os
import subprocess

import numpy as np
import pytest

import _parity as P

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _ndev():
    from ftk_b200 import _lib
    return _lib.lib().ftkb_device_count()


def _device_lists():
    n = _ndev()
    out = [[0, 0], [0, 0, 0]]
    if n >= 2:
        out.append([0, 1])
    if n >= 4:
        out.append([0, 1, 2, 3])
    return out


def _same(a, b, what):
    P.assert_same_result({"points": a.get_discrete_critical_points(), "trajectories": a.get_trajectory_index()},
                         {"points": b.get_discrete_critical_points(), "trajectories": b.get_trajectory_index()}, tol=0.0, what=what)


@pytest.mark.parametrize("chunk", [1, 3, 4, 50])
def test_group_equals_one_device_scalar_2d(chunk, oracle):
    import ftk_b200
    from ftk_b200.group import track_on_devices
    dims, T = [96, 80], 11
    snaps = list(oracle.synthetic_series("woven", dims, T, None))
    one = ftk_b200.track(snaps, dims, field="scalar")
    for ids in _device_lists():
        g, root = track_on_devices(snaps, dims, ids, field="scalar", chunk=chunk)
        _same(root, one, f"group {ids} chunk {chunk}")
        st = g.stats()
        assert st["simplices_tested"] == one.stats()["simplices_tested"]
        g.close()
    one.close()


def test_group_unsaturated_factor_runs_chunks_in_order(oracle):
    """integer-valued fields never saturate the quantisation factor (nbits stays below 21): chunks inherit the exact running
    minimum from their predecessor, one after another -- the factor of every sweep equals the sequential run's"""
    import ftk_b200
    from ftk_b200.group import track_on_devices
    rng = np.random.default_rng(3)
    dims, T = [40, 33], 9
    snaps = [rng.integers(-3, 4, size=(33, 40)).astype(np.float64) * (0.5 if k > 4 else 1.0) for k in range(T)]   # the minimum drops in chunk 2
    one = ftk_b200.track(snaps, dims, field="scalar")
    want = P.oracle_result(oracle.track(snaps, dims, field="scalar"))
    for ids in _device_lists()[:3]:
        g, root = track_on_devices(snaps, dims, ids, field="scalar", chunk=2)
        _same(root, one, f"group {ids} integer field")
        P.assert_same_result({"points": root.get_discrete_critical_points(), "trajectories": root.get_trajectory_index()}, want, tol=TOL, what="vs oracle")
        g.close()
    one.close()


def test_group_3d_vector_and_synthetic(oracle):
    import ftk_b200
    from ftk_b200.group import GroupTracker, track_on_devices
    dims, T = [20, 18, 16], 7
    snaps = list(oracle.synthetic_series("abc", dims, T, None))
    one = ftk_b200.track(snaps, dims, field="vector")
    g, root = track_on_devices(snaps, dims, _device_lists()[-1], field="vector", chunk=3)
    _same(root, one, "group 3D vector")
    g.close()
    one.close()
    # device-side generator: nothing crosses the host, the caller's thread runs ahead of the devices
    dims, T, prm = [64, 48], 13, [30.3, 21.7, 0.4, 0.3]
    one = ftk_b200.make_tracker(dims, field="scalar")
    g = GroupTracker(dims, _device_lists()[-1], field="scalar", chunk=4)
    for tr in (one, g):
        for k in range(T):
            tr.push_synthetic_snapshot(0, prm, float(k))
            if k:
                tr.advance_timestep()
            if k == T - 1:
                tr.update_timestep()
    one.finalize()
    _same(g.finalize(), one, "group synthetic")
    g.close()
    one.close()


def test_cli_device_list_equals_one_device(tmp_path):
    """`ftkb200 -f cp --synthetic woven --device 0,0` (and 0,1 with two GPUs) writes the same file as `--device 0`"""
    from ftk_b200 import build
    cli = build.build_cli()
    outs = []
    lists = ["0", "0,0"] + (["0,1"] if _ndev() >= 2 else [])
    for k, ids in enumerate(lists):
        out = tmp_path / f"traced_{k}.txt"
        cmd = [cli, "-f", "cp", "--synthetic", "woven", "--width", "48", "--height", "40", "--timesteps", "10", "--device", ids, "--time-chunk", "3",
               "-o", str(out)]
        subprocess.run(cmd, check=True, capture_output=True, timeout=300)
        outs.append(out.read_bytes())
    assert len(outs[0]) > 0
    for o in outs[1:]:
        assert o == outs[0]


# ---- spatial decomposition: z-slabs with ghost planes (chunk = 0; the reference's decomposition, regular_tracker.hh:126-149) ------
def _slab_lists():
    n = _ndev()
    out = [[0, 0], [0, 0, 0]]
    if n >= 2:
        out.append([0, 1])
    if n >= 4:
        out.append([0, 1, 2, 3])
    return out


@pytest.mark.parametrize("field,kind", [("scalar", "smooth"), ("vector", "smooth"), ("scalar", "int"), ("vector", "int")])
def test_z_slabs_equal_one_device(field, kind, oracle):
    """every device holds a z-slab (plus ghost planes) of every snapshot and sweeps every step; SoS ranks, positions, corners and
    the quantisation factor are those of the undivided domain, so the merged result equals the one-device run bit for bit"""
    import ftk_b200
    from ftk_b200.group import track_on_devices
    from test_gpu_parity import _rand_series
    rng = np.random.default_rng(91)
    dims, T = [36, 30, 26], 4
    nv = 1 if field == "scalar" else 3
    snaps = _rand_series(rng, dims, T, nv, kind)
    one = ftk_b200.track(snaps, dims, field=field)
    want = P.oracle_result(oracle.track(snaps, dims, field=field))
    assert len(want["points"]) > 0
    for ids in _slab_lists():
        g, root = track_on_devices(snaps, dims, ids, field=field, chunk=0)
        _same(root, one, f"z-slabs {ids} {field} {kind}")
        got = {"points": root.get_discrete_critical_points().copy(), "trajectories": root.get_trajectory_index()}
        if field == "scalar" and kind == "int":
            # 3D types come from an eigenvalue formula through libm (pow / acos / cos, critical_point_type.hh): on an integer-valued
            # field a few thousandths of a per cent of the Jacobians have an eigenvalue that is exactly zero in exact arithmetic, and
            # the last bit of CUDA's vs glibc's libm decides its sign.  Everything else is compared bit for bit.
            differ = got["points"]["cp_type"] != want["points"]["cp_type"]
            assert differ.mean() < 1e-4
            got["points"]["cp_type"] = want["points"]["cp_type"]
        P.assert_same_result(got, want, tol=TOL, what="vs oracle")
        assert root.stats()["points"] >= len(want["points"])      # (the flat simplices of a cut plane are found by both neighbours, merged as one)
        g.close()
    one.close()


def test_z_slabs_moving_extremum_and_generators(oracle):
    """a feature that crosses the cuts (moving extremum through z), and device-side generators that depend on the global z"""
    import ftk_b200
    from ftk_b200.group import GroupTracker, track_on_devices
    dims, T = [32, 28, 40], 6
    x0, d = [15.3, 13.7, 10.1], [0.2, 0.1, 3.7]        # moves 3.7 planes per step: through every cut of 2 and 3 slabs
    snaps = [oracle.gen_moving_extremum(dims, x0, d, float(k)) for k in range(T)]
    one = ftk_b200.track(snaps, dims, field="scalar")
    assert len(one.get_trajectory_index()) == 1
    for ids in _slab_lists():
        g, root = track_on_devices(snaps, dims, ids, field="scalar", chunk=0)
        _same(root, one, f"z-slabs {ids} moving extremum")
        g.close()
    one.close()
    for kind, prm, field, dims in ((3, [np.sqrt(3.0), np.sqrt(2.0), 1.0], "vector", [24, 22, 30]), (5, [], "vector", [20, 18, 24])):
        one = ftk_b200.make_tracker(dims, field=field)
        g = GroupTracker(dims, _slab_lists()[-1], field=field, chunk=0)
        for tr in (one, g):
            for k in range(4):
                tr.push_synthetic_snapshot(kind, prm, float(k))
                if k:
                    tr.advance_timestep()
                if k == 3:
                    tr.update_timestep()
        one.finalize()
        _same(g.finalize(), one, f"z-slabs generator {kind}")
        g.close()
        one.close()
