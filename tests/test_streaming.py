"""Streaming trajectories: trace_critical_points_online, critical_point_tracker.hh:522-641.

tests/golden/stream/<case>.npz hold the trajectories the UNMODIFIED reference grew with
set_enable_streaming_trajectories(true) (made by tests/golden/make_golden_stream.py), as CSR over the sorted points
of tests/golden/<case>.npz, in trajectory-id order with loop / complete flags.  Index work: everything is compared
exactly, ORDER INCLUDED (ids decide who claims a shared neighbour in later steps).
"""
import glob
import os

import numpy as np
import pytest

from _parity import GOLDEN_DIR, golden_snapshots, load_golden

STREAM_DIR = os.path.join(GOLDEN_DIR, "stream")
NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(STREAM_DIR, "*.npz")))


def load_stream(name):
    z = np.load(os.path.join(STREAM_DIR, name + ".npz"))
    off = z["traj_offsets"]
    return [(z["traj_idx"][off[i]:off[i + 1]].astype(np.int64).tolist(), bool(z["traj_loop"][i]), bool(z["traj_complete"][i]))
            for i in range(len(off) - 1)]


@pytest.fixture(scope="module")
def ftkb():
    from ftk_b200 import _lib
    _lib.lib()
    return _lib


def test_fixtures_present():
    assert len(NAMES) >= 15
    assert sum(len(load_stream(n)) for n in NAMES) > 700


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference(name, oracle):
    """the restatement (oracle/cp_online.py) reproduces the reference's streamed trajectories from its points"""
    from oracle import cp_online
    meta, gold, _ = load_golden(name)
    p = gold["points"]
    got = cp_online.trace_streaming(meta["nd"], p["corner"], p["simplex_type"], p["timestep"], meta["T"])
    assert [(list(i), l, c) for i, l, c in got] == load_stream(name)


def _domain(meta):
    margin = 2 if meta["nv"] == 1 else 1          # json_interface.hh:634-656
    return [margin] * meta["nd"], [d - 2 for d in meta["dims"]]


def _index_of(points):
    return {(tuple(int(v) for v in c), int(t)): i for i, (c, t) in enumerate(zip(points["corner"], points["simplex_type"]))}


@pytest.mark.parametrize("prepared", [False, True])
@pytest.mark.parametrize("name", NAMES)
def test_host_grow_step_matches_reference(name, prepared, ftkb):
    """the library's grow step (ftkb_online_*, host code) fed the reference's points one timestep at a time, in a
    scrambled order within the step (the sweep appends hits in no particular order)"""
    from ftk_b200.online import OnlineTracer
    meta, gold, _ = load_golden(name)
    p = np.zeros(len(gold["points"]), ftkb.POINT_DTYPE)
    for f in p.dtype.names:
        p[f] = gold["points"][f]
    lb, ub = _domain(meta)
    tr = OnlineTracer(lb, ub)
    rng = np.random.default_rng(7)
    for j in range(meta["T"] - 1):
        batch = p[p["timestep"] == j]
        tr.grow(batch[rng.permutation(len(batch))], prepared=prepared)
    index = _index_of(p)
    got = [([index[(tuple(int(v) for v in q["corner"]), int(q["simplex_type"]))] for q in pts], l, c) for pts, l, c in tr.trajectories()]
    assert got == load_stream(name)


@pytest.mark.parametrize("prepared", [False, True])
def test_host_grow_step_duplicates_and_empty(prepared, ftkb):
    """an element reported twice in one step is one map entry; empty steps complete every open trajectory"""
    from ftk_b200.online import OnlineTracer
    meta, gold, _ = load_golden("mx2d_11x13x20")
    p = np.zeros(len(gold["points"]), ftkb.POINT_DTYPE)
    for f in p.dtype.names:
        p[f] = gold["points"][f]
    lb, ub = _domain(meta)
    a, b = OnlineTracer(lb, ub), OnlineTracer(lb, ub)
    for j in range(5):
        batch = p[p["timestep"] == j]
        a.grow(batch, prepared=prepared)
        b.grow(np.concatenate([batch, batch[::-1]]), prepared=prepared)
    ta, tb = a.trajectories(), b.trajectories()
    assert len(ta) == len(tb) == 1 and np.array_equal(ta[0][0], tb[0][0]) and not ta[0][2]
    a.grow(p[:0], prepared=prepared)
    assert a.trajectories()[0][2]                       # nothing to add: complete
    a.grow(p[p["timestep"] == 5], prepared=prepared)
    t = a.trajectories()
    assert len(t) == 2 and np.array_equal(t[0][0], ta[0][0])   # a complete trajectory is never extended again


def test_cli_lists_stream_flag():
    import subprocess
    from ftk_b200 import build
    build.build()
    out = subprocess.run([build.CLI, "--help"], capture_output=True, text=True)
    assert out.returncode == 0 and "--stream" in out.stdout


def test_streaming_call_order_is_checked(ftkb):
    import ctypes as C
    assert ftkb.lib().ftkb_set_streaming_trajectories(None, 1) == 1          # FTKB_ERR_INVALID
    h = C.c_void_p()
    assert ftkb.lib().ftkb_online_create(4, None, None, C.byref(h)) != 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_streaming_matches_reference(name, oracle):
    """the tracker with set_enable_streaming_trajectories(True): sweep on the device, grow step after every interval"""
    from ftk_b200 import tracker as T
    meta, gold, inp = load_golden(name)
    field = "scalar" if meta["nv"] == 1 else "vector"
    kw = {} if meta.get("symmetric") is None else {"jacobian_symmetric": bool(meta["symmetric"])}
    tr = T.track(golden_snapshots(meta, inp, oracle), meta["dims"], field=field, streaming=True, **kw)
    pts = tr.get_discrete_critical_points()
    assert np.array_equal(pts["corner"], gold["points"]["corner"]) and np.array_equal(pts["simplex_type"], gold["points"]["simplex_type"])
    complete = tr.get_trajectory_complete()
    got = [(idx.tolist(), loop, bool(complete[i])) for i, (idx, loop) in enumerate(tr.get_trajectory_index())]
    assert got == load_stream(name)
    tr.close()


@pytest.mark.gpu
def test_gpu_streaming_equals_offline_partition_on_simple_case(oracle):
    """moving extremum: one feature, so the streamed trajectory is the offline one minus the last ordinal sweep's point"""
    from ftk_b200 import tracker as T
    snaps = list(oracle.synthetic_series("moving_extremum", [21, 21], 32, None))
    a = T.track(snaps, [21, 21], streaming=True)
    b = T.track(snaps, [21, 21])
    ia, ib = a.get_trajectory_index(), b.get_trajectory_index()
    assert len(ia) == len(ib) == 1
    assert set(ia[0][0].tolist()) <= set(ib[0][0].tolist()) and len(ib[0][0]) - len(ia[0][0]) == 1


@pytest.mark.gpu
def test_cli_stream_flag(tmp_path):
    """`ftk -f cp --stream` (src/cli/ftk.cpp:47,202-203): traced text output holds the streamed trajectories, in id order"""
    import subprocess
    from ftk_b200 import build
    build.build()
    name = "woven_cli_31x37x32"
    want = load_stream(name)
    out = tmp_path / "traced.txt"
    r = subprocess.run([build.CLI, "-f", "cp", "--synthetic", "woven", "--width", "31", "--height", "37", "--timesteps", "32", "--stream",
                        "-o", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    text = open(out).read().splitlines()
    assert text[0] == f"#trajectories={len(want)}"
    lengths, cur = [], None
    for line in text[1:]:
        if line.startswith("--trajectory"):
            if cur is not None:
                lengths.append(cur)
            cur = 0
        elif line.startswith("---"):
            cur += 1
    lengths.append(cur)
    assert lengths == [len(idx) for idx, _, _ in want]


@pytest.mark.parametrize("prepared", [False, True])
@pytest.mark.parametrize("nd,density,seed", [(2, 0.08, 1), (2, 0.3, 2), (2, 0.7, 3), (3, 0.02, 4), (3, 0.1, 5), (3, 0.35, 6)])
def test_host_grow_step_matches_oracle_on_random_sets(nd, density, seed, prepared, ftkb, oracle):
    """random punctured sets, sparse to dense: branching (special) nodes, components of special nodes only, components
    large enough that union_find's doubled sizes wrap around 2^64 -- the library's grow step against the restatement"""
    from oracle import cp_online
    from ftk_b200.online import OnlineTracer
    rng = np.random.default_rng(seed)
    ntypes = 12 if nd == 2 else 60
    side, T = (7, 5) if nd == 2 else (4, 4)
    lb, ub = [0] * nd, [side + 1] * nd
    elems = []
    for t in range(T):
        for idx in np.ndindex(*([side] * nd)):
            for ty in range(ntypes):
                if rng.random() < density:
                    c = [int(v) + 1 for v in idx]                      # away from the border: every vertex in the domain
                    elems.append((c[0], c[1], c[2] if nd == 3 else 0, t, ty))
    assert len(elems) > 50
    p = np.zeros(len(elems), ftkb.POINT_DTYPE)
    p["corner"] = [e[:4] for e in elems]
    p["simplex_type"] = [e[4] for e in elems]
    p["timestep"] = p["corner"][:, 3]
    a, b = cp_online.OnlineTracer(nd), OnlineTracer(lb, ub)
    for t in range(T):
        sel = p[p["timestep"] == t]
        a.grow([tuple(int(v) for v in q["corner"]) + (int(q["simplex_type"]),) for q in sel])
        b.grow(sel[rng.permutation(len(sel))], prepared=prepared)
    want = [(t["elements"], t["loop"], t["complete"]) for t in a.trajectories]
    got = [([tuple(int(v) for v in q["corner"]) + (int(q["simplex_type"]),) for q in pts], l, c) for pts, l, c in b.trajectories()]
    assert got == want
    if density >= 0.3:
        assert sum(len(e) for e, _, _ in want) < len(elems)               # branching nodes were dropped, as in the reference


# ---- time slabs (world_size 2, gloo): streaming=True replays the grow steps on rank 0 ----------------------------------
def _slab_worker(rank, world, port, case, out_path):
    import pickle
    import torch.distributed as dist
    import test_distributed as TD
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        from ftk_b200 import distributed as D
        dims, T, field, snaps = TD.make_series(case)
        tr, info = D.track_time_sharded(lambda k: snaps[k], dims, T, field=field, streaming=True,
                                        tracker_factory=lambda d, f, s, r, **kw: TD.OracleSlabTracker(d, f, s, r, **kw))
        if rank == 0:
            pts = tr.get_discrete_critical_points()
            index = _index_of(pts)
            got = [([index[(tuple(int(v) for v in q["corner"]), int(q["simplex_type"]))] for q in p], l, c) for p, l, c in info["streamed"]]
            with open(out_path, "wb") as f:
                pickle.dump({"points": pts, "streamed": got}, f)
        else:
            assert "streamed" not in info
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["woven", "vector", "3d"])
def test_two_slabs_streaming_equals_sequential_streaming(case, tmp_path, oracle):
    """two time slabs over gloo (the oracle stands in for the sweep): rank 0's replay of the grow steps over the gathered
    punctured simplices == the restatement's streamed trajectories of the sequential run"""
    import pickle
    import torch.multiprocessing as mp
    import test_distributed as TD
    from oracle import cp_online
    out = str(tmp_path / "res")
    mp.spawn(_slab_worker, args=(2, TD._free_port(), case, out), nprocs=2, join=True)
    res = pickle.load(open(out, "rb"))
    dims, T, field, snaps = TD.make_series(case)
    seq = oracle.track(snaps, dims, field=field, trace=False).points()
    assert np.array_equal(res["points"]["corner"], seq["corner"]) and np.array_equal(res["points"]["simplex_type"], seq["simplex_type"])
    want = cp_online.trace_streaming(len(dims), seq["corner"], seq["simplex_type"], seq["timestep"], T)
    assert len(want) > 0 and res["streamed"] == [(list(i), l, c) for i, l, c in want]


def test_host_grow_step_nothing_to_grow(ftkb):
    """no grow step at all (a one-timestep run) and grow steps that receive nothing: no trajectories"""
    from ftk_b200.online import OnlineTracer, replay_streaming
    tr = OnlineTracer([2, 2], [10, 10])
    assert tr.trajectories() == []
    tr.grow(np.zeros(0, ftkb.POINT_DTYPE))
    assert tr.trajectories() == []
    meta, gold, _ = load_golden("mx2d_11x13x20")
    p = np.zeros(len(gold["points"]), ftkb.POINT_DTYPE)
    for f in p.dtype.names:
        p[f] = gold["points"][f]
    assert replay_streaming(p[p["timestep"] == 0], *_domain(meta), T=1) == []      # the only sweep is the last ordinal one
    two = replay_streaming(p[p["timestep"] <= 1], *_domain(meta), T=2)
    assert len(two) == 1 and len(two[0][0]) == int((p["timestep"] == 0).sum())


@pytest.mark.gpu
def test_gpu_streaming_edge_cases(oracle):
    """one timestep (no grow step runs); finalize twice; complete flags reach the curve set; import_points is refused"""
    from ftk_b200 import _lib, tracker as T
    from ftk_b200.curves import CurveSet
    snaps = list(oracle.synthetic_series("woven", [31, 37], 6, None))
    one = T.track(snaps[:1], [31, 37], streaming=True)
    assert one.get_trajectory_index() == [] and len(one.get_discrete_critical_points()) > 0
    one.close()
    tr = T.track(snaps, [31, 37], streaming=True)
    a = [(i.tolist(), l) for i, l in tr.get_trajectory_index()]
    tr.finalize()
    assert [(i.tolist(), l) for i, l in tr.get_trajectory_index()] == a and len(a) > 0
    complete = tr.get_trajectory_complete()
    cs = CurveSet.from_tracker(tr)
    infos, _ = cs.arrays()
    assert [int(v) for v in infos["complete"]] == [int(v) for v in complete] and [int(v) for v in infos["id"]] == list(range(len(a)))
    assert [int(v) for v in infos["count"]] == [len(i) for i, _ in a]
    cs.close()
    pts = tr.get_discrete_critical_points()
    assert _lib.lib().ftkb_import_points(tr._h, pts.ctypes.data, 1) == 1            # FTKB_ERR_INVALID in streaming mode
    assert _lib.lib().ftkb_set_streaming_trajectories(tr._h, 0) == 1                # too late: sweeps have run
    tr.close()


def test_host_grow_step_ignores_foreign_points(ftkb):
    """points outside the domain or with a simplex type the mesh does not have are not elements: ignored, not traced"""
    from ftk_b200.online import OnlineTracer
    p = np.zeros(4, ftkb.POINT_DTYPE)
    p["corner"] = [[3, 3, 0, 0], [50, 3, 0, 0], [3, 3, 0, 0], [3, 3, 0, -1]]
    p["simplex_type"] = [4, 4, 40, 4]                 # 2D+t mesh: 12 types
    tr = OnlineTracer([2, 2], [10, 10])
    tr.grow(p)
    t = tr.trajectories()
    assert len(t) == 1 and len(t[0][0]) == 1 and int(t[0][0]["simplex_type"][0]) == 4
