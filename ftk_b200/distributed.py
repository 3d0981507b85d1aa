"""Time-slab sharding of the tracker over the ranks of a torch.distributed group (one process per GPU).

The path shards along time: the sweeps of different timesteps are independent once the
quantisation factor is known, only the interval sweep of a slab's last timestep needs one layer of the
next slab, and only trace construction needs a global view.  Rank r of R owns the contiguous
timesteps [t0, t1) = slab_range(T, R, r):

  ownership   a simplex belongs to the slab that owns its corner's timestep.  Rank r therefore runs the
              ordinal sweeps at t0 .. t1-1 and the interval sweeps over [t, t+1] for t0 <= t < t1 (the
              last rank has no interval sweep at T-1: the reference's final update_timestep sees one
              snapshot, json_interface.hh:699-706).  Nothing is swept twice, nothing is dropped.
  halo        one layer: the first layer of slab r+1, sent by its owner (send/recv: ncclSend/ncclRecv
              over NVLink for CUDA tensors).  Skipped when the source can produce any layer locally.
              halo="peer" (CUDA tracker, ranks on one node): nothing is sent -- the owner exports CUDA IPC
              handles of the layer and of its 16-byte range cells, the slab below maps them and sweeps the
              layer in place through NVLink peer memory (ftkb_push_snapshot_remote).
  factor      the reference's quantisation factor is a RUNNING quantity (min non-zero |v| over every layer
              seen so far, critical_point_tracker.hh:850-864).  Every rank sweeps optimistically with the
              minimum of its own slab, then the slab minima are all-gathered; a rank whose factor would
              have differed at any of its steps under the exclusive prefix minimum repeats its slab with
              that prefix as the initial resolution (rare: nbits is clamped to [8, 21]).
  merge       the sparse punctured simplices (72-byte records) are gathered on rank 0, which imports them
              and runs the union-find + trace ordering once (ftkb_import_points + ftkb_finalize): the
              labels of components that cross slab boundaries are merged there.

The collective plumbing is backend-agnostic (nccl on the GPUs; gloo in the CPU tests, where the tracker
factory is the test's stand-in); the sweep itself always runs in the CUDA library.
"""
import math

import numpy as np

from . import _lib as L


def slab_range(T, world, rank):
    """contiguous, balanced partition of timesteps 0..T-1: the first T % world slabs get one more"""
    base, extra = divmod(int(T), int(world))
    t0 = rank * base + min(rank, extra)
    return t0, t0 + base + (1 if rank < extra else 0)


def nbits_of(resolution):
    """ref: critical_point_tracker.hh:850-864 (minbits 8, maxbits 21)"""
    if not (resolution > 0) or math.isinf(resolution):
        return 8
    return max(8, min(int(math.ceil(math.log2(1.0 / resolution))), 21))


def exclusive_prefix_min(values, rank):
    return min([float(v) for v in values[:rank]] + [float("inf")])


def _default_factory(dims, field, start_timestep, resolution_init, **kw):
    from .tracker import make_tracker
    return make_tracker(dims, field=field, start_timestep=start_timestep, resolution_init=resolution_init, **kw)


def _sweep_slab(make, layer, push, t0, t1, T, halo_layer, resolution_init):
    """one pass over the slab; returns (tracker, [(running resolution, factor) per sweep])"""
    tr = make(t0, resolution_init)
    log = []
    last = t1 == T
    for k in range(t0, t1 + (0 if last else 1)):
        push(tr, halo_layer if (k == t1 and halo_layer is not None) else layer(k))
        if k > t0:
            tr.advance_timestep()
            log.append(_res_factor(tr))
    if last:
        tr.update_timestep()
        log.append(_res_factor(tr))
    return tr, log


def _res_factor(tr):
    st = tr.stats()
    return float(st["resolution"]), float(st["scaling_factor"])


def _peer_sweep_slab(make, layer, push, t0, t1, T, field, resolution_init, exchange, keep):
    """one pass over the slab with the halo read in place from the next rank (see track_time_sharded, halo="peer").
    exchange(mine) -> handles of the next rank's first layer (or None for the last rank); called once, after this rank's
    first sweep has built the range cells of its own first layer."""
    import torch
    tr = make(t0, resolution_init)
    log = []
    last = t1 == T
    first = layer(t0)
    if not torch.is_tensor(first) or not first.is_cuda:
        first = torch.from_numpy(np.ascontiguousarray(first, dtype=np.float64)).cuda()
    first = first.contiguous()
    keep.append(first)                                     # the neighbour reads it in place: must outlive the sweep
    kw = "scalar" if field == "scalar" else "vector"
    tr.push_device_pointers(**{kw: int(first.data_ptr())})
    tr.last_layer_resolution()                             # streams the first layer once: its range cells and min |v|, no sweep
    cptr, _, cres = tr.export_layer_cells(0)
    peer = exchange((L.ipc_export(int(first.data_ptr())), L.ipc_export(cptr), cres))
    push(tr, layer(t0 + 1))
    tr.advance_timestep()                                  # every step of the slab is swept exactly once
    log.append(_res_factor(tr))
    for k in range(t0 + 2, t1):
        push(tr, layer(k))
        tr.advance_timestep()
        log.append(_res_factor(tr))
    if last:
        tr.update_timestep()
    else:
        import torch.cuda
        dev = torch.cuda.current_device()
        mapped = [(L.ipc_import(peer[0], dev), peer[0]), (L.ipc_import(peer[1], dev), peer[1])]
        tr.push_remote_snapshot(**{kw: mapped[0][0], "cells": mapped[1][0], "resolution": peer[2]})
        tr.advance_timestep()
        tr.synchronize()
        keep.append(("ipc", mapped))                       # unmapped once nobody sweeps any more (after the barrier)
    log.append(_res_factor(tr))
    return tr, log


def _streamed(tr, dims, field, T, kw):
    """streaming=True on time slabs: rank 0 replays the reference's grow steps over the gathered punctured simplices"""
    from .online import replay_streaming
    margin = 2 if field == "scalar" else 1           # make_tracker's default domain (json_interface.hh:634-656)
    lb = kw.get("lb") or [margin] * len(dims)
    ub = kw.get("ub") or [d - 2 for d in dims]
    return replay_streaming(tr.get_discrete_critical_points(), lb, ub, T)


def track_time_sharded(layer, dims, T, field="scalar", group=None, tracker_factory=None, exchange_halo=True, trace=True, halo="copy",
                       streaming=False, **kw):
    """Run the tracker over T timesteps sharded into time slabs over the ranks of `group`.

    layer(k)         -> snapshot k in memory order (numpy array or CUDA tensor); called on the rank that owns
                        timestep k (and for the halo layer k = t1 when exchange_halo is False)
    tracker_factory  (dims, field, start_timestep, resolution_init, **kw) -> tracker; default: the CUDA tracker
    halo             "copy": the halo layer is sent (send/recv); "peer": it is read in place through NVLink peer memory
                     (CUDA tracker only, every slab at least two timesteps, all ranks on one node)
    streaming        True: instead of the offline trace, rank 0 replays the reference's streaming grow steps
                     (trace_critical_points_online, critical_point_tracker.hh:522-641) over the gathered punctured simplices,
                     timestep by timestep; the trajectories [(points, loop, complete)] are returned in info["streamed"]
    Returns the rank-0 tracker holding every punctured simplex (finalized when trace=True) on rank 0 and the
    local slab's tracker elsewhere, plus a dict of bookkeeping (slab, halo bytes, repeated slabs)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if T < world:
        raise ValueError("fewer timesteps than ranks")
    t0, t1 = slab_range(T, world, rank)
    factory = tracker_factory or _default_factory

    def make(start, res_init):
        return factory(dims, field, start, res_init, **kw)

    def push(tr, a):
        if field == "scalar":
            tr.push_scalar_field_snapshot(a)
        else:
            tr.push_vector_field_snapshot(a)

    info = {"rank": rank, "world": world, "slab": (t0, t1), "halo_bytes": 0, "slab_repeated": False, "halo": "copy"}
    peer_mode = halo == "peer" and world > 1 and tracker_factory is None and all(b - a >= 2 for a, b in (slab_range(T, world, r) for r in range(world)))
    if peer_mode:
        info["halo"] = "peer"
        keep, handles = [], {}

        def exchange(mine):
            # every rank contributes the handles of its first layer; a rank uses those of the rank above
            if "all" not in handles:
                allh = [None] * world
                dist.all_gather_object(allh, mine, group=group)
                handles["all"] = allh
            return handles["all"][rank + 1] if rank < world - 1 else None

        tr, log = _peer_sweep_slab(make, layer, push, t0, t1, T, field, 0.0, exchange, keep)
        mine = torch.tensor([min(r for r, _ in log)], dtype=torch.float64)
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        mine = mine.to(dev)
        allres = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allres, mine, group=group)
        prefix = exclusive_prefix_min([float(a.item()) for a in allres], rank)
        info["prefix_resolution"] = prefix
        if prefix < float("inf") and any(float(1 << nbits_of(min(prefix, r))) != f for r, f in log):
            # repeat with the inherited minimum; the first pass's tracker stays alive: the rank below may still read its cells
            keep.append(tr)
            tr, log = _peer_sweep_slab(make, layer, push, t0, t1, T, field, prefix, exchange, keep)
            info["slab_repeated"] = True
        info["factors"] = [f for _, f in log]
        pts = tr.get_discrete_critical_points()
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(pts.tobytes(), gathered, dst=0, group=group)
        if rank == 0:
            for b in gathered[1:]:
                if len(b):
                    tr.import_points(np.frombuffer(b, dtype=L.POINT_DTYPE))
        dist.barrier(group=group)             # nobody releases a layer or its cells while a neighbour may still sweep them
        for t in keep:
            if isinstance(t, tuple) and t and t[0] == "ipc":
                for ptr, handle in t[1]:
                    try:
                        L.ipc_close(ptr, handle)
                    except L.FTKBError:
                        pass
            elif hasattr(t, "close"):
                t.close()
        if rank == 0 and streaming:
            info["streamed"] = _streamed(tr, dims, field, T, kw)
        elif rank == 0 and trace:
            tr.finalize()
        return tr, info

    # ---- halo: first layer of the next slab --------------------------------------------------------------
    halo = None
    first = None
    if world > 1 and exchange_halo:
        reqs = []
        if rank > 0:
            first = layer(t0)
            send_t = first if torch.is_tensor(first) else torch.from_numpy(np.ascontiguousarray(first, dtype=np.float64))
            reqs.append(dist.isend(send_t.contiguous(), rank - 1, group=group))
        if rank < world - 1:
            proto = first if first is not None else layer(t0)
            if first is None:
                first = proto
            halo = torch.empty_like(proto) if torch.is_tensor(proto) else torch.empty(np.shape(proto), dtype=torch.float64)
            reqs.append(dist.irecv(halo, rank + 1, group=group))
            info["halo_bytes"] = halo.numel() * 8
        for r in reqs:
            r.wait()
        if halo is not None and halo.is_cuda:
            torch.cuda.current_stream(halo.device).synchronize()
        if halo is not None and not halo.is_cuda:
            halo = halo.numpy()
    cached_first = first

    def layer_c(k):
        return cached_first if (k == t0 and cached_first is not None) else layer(k)

    # ---- optimistic sweep with the slab's own running minimum ---------------------------------------------
    tr, log = _sweep_slab(make, layer_c, push, t0, t1, T, halo, 0.0)
    if world > 1:
        mine = torch.tensor([min(r for r, _ in log)], dtype=torch.float64)
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        mine = mine.to(dev)
        allres = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allres, mine, group=group)
        prefix = exclusive_prefix_min([float(a.item()) for a in allres], rank)
        info["prefix_resolution"] = prefix
        if prefix < float("inf") and any(float(1 << nbits_of(min(prefix, r))) != f for r, f in log):
            # the factor this slab used differs from the sequential run's: repeat it with the inherited minimum
            tr.close() if hasattr(tr, "close") else None
            tr, log = _sweep_slab(make, layer_c, push, t0, t1, T, halo, prefix)
            info["slab_repeated"] = True
    info["factors"] = [f for _, f in log]

    # ---- merge on rank 0 -------------------------------------------------------------------------------------
    pts = tr.get_discrete_critical_points()
    if world > 1:
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(pts.tobytes(), gathered, dst=0, group=group)
        if rank == 0:
            for b in gathered[1:]:
                if len(b):
                    tr.import_points(np.frombuffer(b, dtype=L.POINT_DTYPE))
    if rank == 0 and streaming:
        info["streamed"] = _streamed(tr, dims, field, T, kw)
    elif rank == 0 and trace:
        tr.finalize()
    return tr, info
