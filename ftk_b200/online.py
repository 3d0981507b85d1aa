"""The streaming grow step by itself (host code of libftkb200.so, no device needed):
trace_critical_points_online, ref include/ftk/filters/critical_point_tracker.hh:522-641.

A tracker created with set_enable_streaming_trajectories(True) runs this after every interval sweep; this class
lets a caller that already holds punctured simplices (read from an archive, gathered from time slabs) grow the
same trajectories step by step.
"""
import ctypes as C

import numpy as np

from . import _lib as L


class OnlineTracer:
    def __init__(self, lb, ub):
        """lb, ub: tracker domain, inclusive, 2 or 3 entries (set_domain, regular_tracker.hh:24)"""
        nd = len(lb)
        a, b = (C.c_int32 * nd)(*[int(v) for v in lb]), (C.c_int32 * nd)(*[int(v) for v in ub])
        h = C.c_void_p()
        rc = L.lib().ftkb_online_create(nd, a, b, C.byref(h))
        if rc:
            raise L.FTKBError(rc, "ftkb_online_create failed")
        self._h = h

    def close(self):
        if self._h:
            L.lib().ftkb_online_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def grow(self, pts, prepared=False):
        """pts: structured array (L.POINT_DTYPE) of the punctured simplices found since the last call; prepared: walk over
        sorted neighbour lists (what the tracker's device-side preparation produces) instead of a hash of the batch"""
        pts = np.ascontiguousarray(pts, dtype=L.POINT_DTYPE)
        f = L.lib().ftkb_online_grow_prepared if prepared else L.lib().ftkb_online_grow
        rc = f(self._h, pts.ctypes.data, len(pts))
        if rc:
            raise L.FTKBError(rc, "ftkb_online_grow failed")

    def trajectories(self):
        """-> list of (points in trace order, loop, complete), in trajectory-id order"""
        nt, npt = C.c_uint64(), C.c_uint64()
        L.lib().ftkb_online_size(self._h, C.byref(nt), C.byref(npt))
        off = np.zeros(nt.value + 1, np.uint64)
        pts = np.zeros(max(npt.value, 1), L.POINT_DTYPE)
        loop, complete = np.zeros(max(nt.value, 1), np.uint8), np.zeros(max(nt.value, 1), np.uint8)
        L.lib().ftkb_online_get(self._h, off.ctypes.data, pts.ctypes.data, loop.ctypes.data, complete.ctypes.data)
        return [(pts[int(off[i]):int(off[i + 1])], bool(loop[i]), bool(complete[i])) for i in range(nt.value)]


def replay_streaming(points, lb, ub, T):
    """Streaming trajectories of a finished run: the grow steps the reference runs after the sweeps of timesteps
    0 .. T-2 (critical_point_tracker_2d_regular.hh:322-330; the last ordinal sweep, timestep T-1, is followed by none),
    replayed over `points` (L.POINT_DTYPE, any order).  This is how the time-slab driver (distributed.py) serves
    streaming=True: rank 0 holds every slab's punctured simplices after the gather, and the grow steps only depend on
    which timestep's sweep found a point.  -> [(points in trace order, loop, complete)] in trajectory-id order."""
    points = np.ascontiguousarray(points, dtype=L.POINT_DTYPE)
    tr = OnlineTracer(lb, ub)
    order = np.argsort(points["timestep"], kind="stable")
    ts = points["timestep"][order]
    for j in range(int(T) - 1):
        a, b = np.searchsorted(ts, j, "left"), np.searchsorted(ts, j, "right")
        tr.grow(points[order[a:b]])
    out = tr.trajectories()
    tr.close()
    return out
