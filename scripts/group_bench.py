#!/usr/bin/env python
"""One tracker over several GPUs of one process (ftkb_group): ms per step of the 512^3 scalar configuration (C3's
generator) with z-slabs, (a) snapshots generated on the devices, (b) host snapshots in page-locked memory pushed through
ftkb_group_push_snapshot -- every device copies its own slab over its own PCIe link.
    python scripts/group_bench.py [--dims 512 512 512] [--steps 12] > gpurun_out/group_bench.jsonl"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ftk_b200 import _lib as L  # noqa: E402
from ftk_b200.group import GroupTracker  # noqa: E402
import ftk_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dims", type=int, nargs=3, default=[512, 512, 512])
ap.add_argument("--steps", type=int, default=12)
args = ap.parse_args()
dims, K = args.dims, args.steps
ndev = L.lib().ftkb_device_count()
x0, d = [dims[0] / 2 + 0.3, dims[1] / 2 - 0.3, dims[2] / 2 + 0.1], [0.1, 0.11, 0.1]
prm = x0 + d
per_step = (dims[0] - 3) * (dims[1] - 3) * (dims[2] - 3) * 60
nvert = dims[0] * dims[1] * dims[2]


def pinned(n):
    p = C.c_void_p()
    assert L.lib().ftkb_host_alloc(n * 8, C.byref(p)) == 0
    return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(n,)), p


def host_layers():
    z, y, x = np.meshgrid(*[np.arange(n, dtype=np.float64) for n in reversed(dims)], indexing="ij", sparse=True)
    out = []
    for k in range(2):
        a, p = pinned(nvert)
        a.reshape(tuple(reversed(dims)))[...] = (x - (x0[0] + d[0] * k)) ** 2 + (y - (x0[1] + d[1] * k)) ** 2 + (z - (x0[2] + d[2] * k)) ** 2
        out.append((a, p))
    return out


def run(ids, mode, layers):
    one = len(ids) == 1
    tr = ftk_b200.make_tracker(dims, field="scalar", device=ids[0]) if one else GroupTracker(dims, ids, field="scalar", chunk=0)

    def push(k):
        if mode == "synthetic":
            tr.push_synthetic_snapshot(0, prm, float(k))
        else:
            tr.push_scalar_field_snapshot(layers[k % 2][0].reshape(tuple(reversed(dims))))
    push(0)
    for k in range(1, 4):                      # warm-up (buffers, first layers, the factor saturates)
        push(k)
        tr.advance_timestep()
    if one:
        tr.synchronize()
    else:
        tr.stats()
    t0 = time.perf_counter()
    for k in range(4, 4 + K):
        push(k)
        tr.advance_timestep()
    root = tr if one else tr.finalize()        # the group is drained by its finalize (merge of a handful of points)
    if one:
        tr.synchronize()
    dt = time.perf_counter() - t0
    npts = len(root.get_discrete_critical_points())
    tr.close()
    return {"devices": ids, "mode": mode, "dims": dims, "steps": K, "ms_per_step": 1e3 * dt / K, "simplices_per_s": per_step * K / dt,
            "punctured_simplices": npts, "timing": "host wall clock around K push + advance calls (+ drain)"}


layers = host_layers()
for n in (1, 2, 4, 8):
    if n > ndev:
        break
    for mode in ("synthetic", "host"):
        print(json.dumps(run(list(range(n)), mode, layers)), flush=True)
