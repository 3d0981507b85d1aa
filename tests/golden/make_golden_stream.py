"""Generate tests/golden/stream/*.npz: trajectories of the UNMODIFIED reference run with streaming trajectories
(set_enable_streaming_trajectories(true): trace_critical_points_online after every interval sweep,
critical_point_tracker.hh:522-641), through oracle/_ref/ftk_ref_oracle --out-stream.

Run in the build container (needs /root/reference for oracle/build_ref.sh):
    python tests/golden/make_golden_stream.py
Each fixture refers to the tracking fixture of the same name (tests/golden/<name>.npz, same case, same sorted points):
it holds the streamed trajectories in trajectory-id order as CSR over those points, with the loop / complete flags.
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cp_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "stream")

CASES = ["woven_10x10x20", "woven_128x128x10", "mx2d_11x13x20", "mx2d_21x21x32", "woven_cli_31x37x32", "merger_32x32x100",
         "double_gyre_64x32x50", "mx3d_21x21x21x10", "abc_24x24x24x4", "tornado_20x20x20x4", "cos3d_14x14x14x4",
         "rand2d_scalar_int", "rand2d_vector_normal_sym", "rand3d_scalar_normal", "rand3d_vector_int"]


def read_stream(path):
    a = np.fromfile(path, np.uint64)
    nt, p = int(a[0]), 1
    ids, off, idx, loop, complete = [], [0], [], [], []
    for _ in range(nt):
        k, n, l, c = (int(v) for v in a[p:p + 4])
        p += 4
        ids.append(k); loop.append(l); complete.append(c)
        idx.append(a[p:p + n].astype(np.int64))
        p += n
        off.append(off[-1] + n)
    assert p == len(a)
    assert ids == list(range(nt)), "trajectory ids are not 0..n-1 in order"
    return (np.asarray(off, np.int64), np.concatenate(idx) if idx else np.zeros(0, np.int64),
            np.asarray(loop, np.uint8), np.asarray(complete, np.uint8))


def main():
    O.build()
    os.makedirs(OUT, exist_ok=True)
    for name in CASES:
        g = np.load(os.path.join(HERE, name + ".npz"))
        meta = json.loads(bytes(g["meta"]).decode())
        cmd = [O.REF_BINARY, "--nd", str(meta["nd"]), "--nv", str(meta["nv"]), "--dims"] + [str(d) for d in meta["dims"]] + \
              ["--nt", str(meta["T"]), "--quiet"]
        tmp = tempfile.mkdtemp()
        if "input" in g.files:
            inp = os.path.join(tmp, "in.f64")
            np.ascontiguousarray(g["input"], np.float64).tofile(inp)
            cmd += ["--input", inp]
        else:
            cmd += ["--gen", meta["gen"]]
            if meta.get("params"):
                cmd += ["--p"] + [repr(float(x)) for x in meta["params"]]
        if meta.get("symmetric") is not None:
            cmd += ["--symmetric", str(int(meta["symmetric"]))]
        out, strm = os.path.join(tmp, "o.ftkg"), os.path.join(tmp, "o.strm")
        cmd += ["--out", out, "--out-stream", strm]
        subprocess.run(cmd, check=True, capture_output=True)
        gold = O.read_ftkg(out)
        pts = gold["points"]
        # the same run must reproduce the tracking fixture's points (the stream indices refer to them)
        assert np.array_equal(pts["corner"], g["corner"]) and np.array_equal(pts["simplex_type"].astype(np.int8), g["simplex_type"]), name
        off, idx, loop, complete = read_stream(strm)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), traj_offsets=off, traj_idx=idx.astype(np.int32), traj_loop=loop,
                            traj_complete=complete)
        print(f"{name}: {len(pts)} points, {len(loop)} streamed trajectories over {len(idx)} points, "
              f"{int(loop.sum())} loops, {int(complete.sum())} complete")


if __name__ == "__main__":
    main()
