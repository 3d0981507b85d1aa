"""Multi-rank time-slab protocol (ftk_b200/distributed.py) with world_size 2 over gloo.

CPU tests plug the parity oracle in as the per-slab tracker (test stand-in: the product's factory is the
CUDA tracker), so they check the host-side protocol -- ownership, halo, running-resolution prefix with the
repeated slab, merge + trace on rank 0 -- against ONE sequential oracle run.  The `gpu` test runs the same
protocol with the CUDA tracker on both ranks (gloo plumbing, both ranks on cuda:0).
"""
import os
import pickle
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class OracleSlabTracker:
    """the oracle behind the tracker interface distributed.py drives (test stand-in)"""

    def __init__(self, dims, field, start_timestep, resolution_init, **kw):
        from oracle import cp_oracle as O
        self.tr = O.Tracker(dims, field=field, start_timestep=start_timestep, **kw)
        if resolution_init > 0:
            self.tr.set_resolution(resolution_init)

    def push_scalar_field_snapshot(self, a):
        self.tr.push_scalar_field_snapshot(a)

    def push_vector_field_snapshot(self, a):
        self.tr.push_vector_field_snapshot(a)

    def advance_timestep(self):
        self.tr.advance_timestep()

    def update_timestep(self):
        self.tr.update_timestep()

    def stats(self):
        return {"resolution": self.tr.resolution, "scaling_factor": self.tr.scaling_factor}

    def get_discrete_critical_points(self):
        return self.tr.points()

    def import_points(self, pts):
        self.tr.import_points(pts)

    def finalize(self):
        self.tr.finalize()

    def get_trajectory_index(self):
        return self.tr.trajectories()

    def close(self):
        pass


def make_series(case):
    rng = np.random.default_rng(5)
    if case == "woven":
        from oracle import cp_oracle as O
        dims, T = [24, 20], 9
        return dims, T, "scalar", list(O.synthetic_series("woven", dims, T))
    if case == "vector":
        from oracle import cp_oracle as O
        dims, T = [32, 16], 7
        return dims, T, "vector", list(O.synthetic_series("double_gyre", dims, T))
    if case == "resolution_prefix":
        # slab 0 holds a tiny non-zero gradient (nbits 20), slab 1 only coarse values (nbits 8 on its own):
        # the second slab must be repeated with the inherited running minimum to match the sequential run
        dims, T = [16, 14], 6
        snaps = [np.round(rng.normal(size=(14, 16)) * 4) * 0.5 for _ in range(T)]
        snaps[1][3:8, 4:11] = 1.0
        snaps[1][5, 7] += 1e-7
        return dims, T, "scalar", snaps
    if case == "3d":
        dims, T = [10, 9, 8], 5
        return dims, T, "scalar", [rng.normal(size=(8, 9, 10)) for _ in range(T)]
    raise ValueError(case)


def _worker(rank, world, port, case, use_cuda, out_path, halo="copy"):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        from ftk_b200 import distributed as D
        dims, T, field, snaps = make_series(case)
        factory = None if use_cuda else (lambda d, f, s, r, **kw: OracleSlabTracker(d, f, s, r, **kw))
        calls = []
        extra = {}
        if use_cuda:
            # one GPU per rank when the box has them (the halo then really crosses NVLink); both ranks on cuda:0 otherwise
            import torch
            dev = rank % max(1, torch.cuda.device_count())
            torch.cuda.set_device(dev)
            extra["device"] = dev

        def layer(k):
            calls.append(k)
            return snaps[k]
        tr, info = D.track_time_sharded(layer, dims, T, field=field, tracker_factory=factory, halo=halo, **extra)
        t0, t1 = D.slab_range(T, world, rank)
        assert info["slab"] == (t0, t1)
        assert all(t0 <= k < t1 for k in calls), (calls, t0, t1)      # a rank only ever asks for its own layers: the halo is exchanged
        res = {"info": {k: v for k, v in info.items()}}
        if rank == 0:
            res["points"] = tr.get_discrete_critical_points()
            res["trajectories"] = tr.get_trajectory_index()
        with open(f"{out_path}.{rank}", "wb") as f:
            pickle.dump(res, f)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _run(case, tmp_path, use_cuda=False, world=2, halo="copy"):
    import torch.multiprocessing as mp
    out = str(tmp_path / "res")
    mp.spawn(_worker, args=(world, _free_port(), case, use_cuda, out, halo), nprocs=world, join=True)
    return [pickle.load(open(f"{out}.{r}", "rb")) for r in range(world)]


def _sequential(case, oracle):
    dims, T, field, snaps = make_series(case)
    o = oracle.track(snaps, dims, field=field)
    return {"points": o.points(), "trajectories": o.trajectories()}, [o.scaling_factor]


def test_slab_range_partitions_time():
    from ftk_b200.distributed import slab_range
    for T in (1, 2, 7, 64, 65):
        for world in (1, 2, 3, 4, 8):
            if T < world:
                continue
            slabs = [slab_range(T, world, r) for r in range(world)]
            assert slabs[0][0] == 0 and slabs[-1][1] == T
            assert all(a[1] == b[0] for a, b in zip(slabs, slabs[1:]))
            sizes = [b - a for a, b in slabs]
            assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 1


def test_nbits_matches_reference_clamp():
    from ftk_b200.distributed import nbits_of, exclusive_prefix_min
    assert nbits_of(1.0) == 8 and nbits_of(1e-30) == 21 and nbits_of(float("inf")) == 8
    assert nbits_of(2.0 ** -10) == 10 and nbits_of(2.0 ** -10 * 1.01) == 10 and nbits_of(2.0 ** -10 * 0.99) == 11
    assert exclusive_prefix_min([3.0, 1.0, 2.0], 0) == float("inf")
    assert exclusive_prefix_min([3.0, 1.0, 2.0], 2) == 1.0


@pytest.mark.parametrize("case", ["woven", "vector", "3d"])
def test_two_slabs_equal_sequential_run(case, tmp_path, oracle):
    import _parity as P
    want, _ = _sequential(case, oracle)
    res = _run(case, tmp_path)
    P.assert_same_result(res[0], want, tol=0.0, what=f"time slabs x2 ({case})")
    assert res[1]["info"]["halo_bytes"] == 0 and res[0]["info"]["halo_bytes"] > 0


def test_running_resolution_crosses_slabs(tmp_path, oracle):
    """the quantisation factor is a running quantity: the second slab has to inherit the first slab's minimum"""
    import _parity as P
    want, _ = _sequential("resolution_prefix", oracle)
    res = _run("resolution_prefix", tmp_path)
    assert res[1]["info"]["slab_repeated"] and not res[0]["info"]["slab_repeated"]
    assert set(res[1]["info"]["factors"]) == {float(1 << 20)}      # 1e-7 * (W - 1) = 1.5e-6 -> 20 bits
    P.assert_same_result(res[0], want, tol=0.0, what="time slabs x2 (inherited resolution)")


def test_three_slabs(tmp_path, oracle):
    import _parity as P
    want, _ = _sequential("woven", oracle)
    res = _run("woven", tmp_path, world=3)
    P.assert_same_result(res[0], want, tol=0.0, what="time slabs x3")


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["woven", "resolution_prefix", "3d"])
def test_two_slabs_cuda(case, tmp_path, oracle):
    """same protocol, CUDA tracker on both ranks (both on cuda:0; gloo carries the halo and the merge)"""
    import _parity as P
    want, _ = _sequential(case, oracle)
    res = _run(case, tmp_path, use_cuda=True)
    P.assert_same_result(res[0], want, tol=1e-9, what=f"CUDA time slabs x2 ({case})")


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["woven", "resolution_prefix", "3d", "vector"])
def test_two_slabs_cuda_peer_memory_halo(case, tmp_path, oracle):
    """halo="peer": the upper slab's first layer is never sent; the lower slab maps it and its range cells (CUDA IPC, NVLink
    peer memory between GPUs) and sweeps them in place -- including the slab that is repeated with an inherited resolution"""
    import _parity as P
    want, _ = _sequential(case, oracle)
    res = _run(case, tmp_path, use_cuda=True, halo="peer")
    assert all(r["info"]["halo"] == "peer" and r["info"]["halo_bytes"] == 0 for r in res)
    if case == "resolution_prefix":
        assert res[1]["info"]["slab_repeated"]
    P.assert_same_result(res[0], want, tol=1e-9, what=f"CUDA time slabs x2, peer-memory halo ({case})")
