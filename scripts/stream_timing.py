"""Cost of streaming trajectories next to the offline trace, on one GPU.

    python scripts/stream_timing.py [W H T]

Woven 2D scalar field synthesised on the device; every step is push + advance_timestep; wall clock around the whole
loop plus finalize (the grow step is host work between sweeps, so device events alone would not see it).  Prints one
JSON line per mode.  Not a bench.py number: bench.py's metric stays the offline path.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from ftk_b200 import _lib as L
from ftk_b200 import tracker as T


def run(W, H, steps, streaming):
    tr = T.make_tracker([W, H], field="scalar", streaming=streaming)
    t0 = time.perf_counter()
    for k in range(steps):
        tr.push_synthetic_snapshot(L.SYN_WOVEN, [], k / (steps - 1))
        if k != 0:
            tr.advance_timestep()
        if k == steps - 1:
            tr.update_timestep()
    tr.synchronize()
    t1 = time.perf_counter()
    tr.finalize()
    t2 = time.perf_counter()
    st = tr.stats()
    ntraj = len(tr.get_trajectory_index())
    out = {"mode": "streaming" if streaming else "offline", "workload": f"woven {W}x{H}x{steps}", "ms_steps": (t1 - t0) * 1e3,
           "ms_per_timestep": (t1 - t0) * 1e3 / steps, "ms_finalize": (t2 - t1) * 1e3, "punctured": int(st["points"]), "trajectories": ntraj,
           "ms_scan": st["ms_scan"], "ms_test": st["ms_test"], "ms_host_trace": st["ms_finalize_host"], "ms_device_trace": st["ms_finalize_device"],
           "d2h_bytes": int(st["d2h_bytes"]), "kernel_launches": int(st["kernel_launches"])}
    tr.close()
    return out


if __name__ == "__main__":
    W, H, steps = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (4096, 4096, 32)
    run(256, 256, 4, False)   # warm-up: context, module load
    for mode in (False, True, False, True):
        print(json.dumps(run(W, H, steps, mode)))
