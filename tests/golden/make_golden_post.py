"""Generate tests/golden/post/*.npz: trajectory post-processing by the UNMODIFIED reference.

    python tests/golden/make_golden_post.py        (build container: needs /root/reference for oracle/build_ref.sh)

oracle/_ref/ftk_ref_oracle tracks a case, then applies --post OPS with the reference's own feature_curve_t /
feature_curve_set_t / feature_curve_set_post_processor_t code and dumps the curve set (--out-curves).  A fixture
holds, per op list, the reference's trajectories of THAT run in trace order (curve ids are positions in that order; the
reference's trace order is not deterministic from run to run -- its components are traced by a thread pool) and the
resulting curves: id / loop / complete / consistent_type / statistics and per point (index into the case's
punctured simplices, type, ordinal, timestep, id, t, v).  The punctured simplices themselves are those of the
tracking fixture tests/golden/<case>.npz.
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import cp_oracle as O  # noqa: E402
import _parity as P  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "post")

CASES = ["merger_32x32x100", "double_gyre_64x32x50", "mx3d_21x21x21x10", "rand2d_scalar_int", "cos3d_14x14x14x4",
         "woven_10x10x20"]
OPS = [
    "smooth_types,rotate,split,reorder,adjust_time",       # feature_curve_set_post_processor_t ops in the legacy order
    "legacy",                                              # json_interface::post_process() defaults
    "legacy:0:1:1",                                        # ... discarding interval points, deriving velocities
    "legacy:2.5",                                          # ... duration pruning
    "split",                                               # finalize() left statistics: only curves of mixed type are split
    "update_statistics,split,discard_interval_points",
    "derive_velocity",
    "update_statistics,discard_degenerate_points,reorder",
    "update_statistics,duration_pruning:4,rotate,smooth_types",
    "reorder,intercept:1:3",                               # feature_curve_set_t::intercept (curves assumed reordered)
]

CURVE_REC = np.dtype([("id", np.int32), ("loop", np.int32), ("complete", np.int32), ("consistent_type", np.uint32), ("count", np.uint64),
                      ("tmin", np.float64), ("tmax", np.float64), ("bbmin", np.float64, 3), ("bbmax", np.float64, 3),
                      ("smin", np.float64), ("smax", np.float64), ("persistence", np.float64), ("vmmin", np.float64), ("vmmax", np.float64)])
POINT_REC = np.dtype([("idx", np.uint64), ("cp_type", np.uint32), ("ordinal", np.int32), ("timestep", np.int32), ("id", np.int32),
                      ("x", np.float64, 3), ("t", np.float64), ("scalar", np.float64), ("v", np.float64, 3)])
assert CURVE_REC.itemsize == 128 and POINT_REC.itemsize == 88


def read_ftkc(path):
    buf = open(path, "rb").read()
    magic, version = np.frombuffer(buf, np.uint32, 2, 0)
    assert magic == 0x434B5446 and version == 1
    nc = int(np.frombuffer(buf, np.uint64, 1, 8)[0])
    off = 16
    curves, points = np.zeros(nc, CURVE_REC), []
    for i in range(nc):
        curves[i] = np.frombuffer(buf, CURVE_REC, 1, off)[0]
        off += CURVE_REC.itemsize
        n = int(curves[i]["count"])
        points.append(np.frombuffer(buf, POINT_REC, n, off).copy())
        off += POINT_REC.itemsize * n
    ns = int(np.frombuffer(buf, np.uint64, 1, off)[0])
    off += 8
    slices = []
    for _ in range(ns):
        t = int(np.frombuffer(buf, np.int32, 1, off)[0])
        n = int(np.frombuffer(buf, np.uint64, 1, off + 8)[0])
        off += 16
        slices.append((t, np.frombuffer(buf, np.uint64, n, off).astype(np.int64)))
        off += 8 * n
    assert off == len(buf)
    return curves, (np.concatenate(points) if points else np.zeros(0, POINT_REC)), slices


def run(meta, inp, ops, tmpdir):
    cmd = [O.REF_BINARY, "--nd", str(meta["nd"]), "--nv", str(meta["nv"]), "--dims"] + [str(d) for d in meta["dims"]] + ["--nt", str(meta["T"]), "--quiet"]
    if inp is not None:
        raw = os.path.join(tmpdir, "in.f64")
        np.ascontiguousarray(inp, np.float64).tofile(raw)
        cmd += ["--input", raw]
    else:
        cmd += ["--gen", meta["gen"]]
        if meta["params"]:
            cmd += ["--p"] + [repr(float(x)) for x in meta["params"]]
    if meta["symmetric"] is not None:
        cmd += ["--symmetric", str(int(meta["symmetric"]))]
    g, c = os.path.join(tmpdir, "o.ftkg"), os.path.join(tmpdir, "o.ftkc")
    subprocess.run(cmd + ["--out", g, "--post", ops, "--out-curves", c], check=True, capture_output=True)
    return O.read_ftkg(g), read_ftkc(c)


def main():
    O.build()
    os.makedirs(OUT, exist_ok=True)
    for name in CASES:
        meta, gold, inp = P.load_golden(name)
        arrays = {"meta": np.frombuffer(json.dumps(dict(case=name, ops=OPS, reference="hguo/ftk@aa4f2cf9 feature_curve(_set)_t, g++ -O2")).encode(), np.uint8)}
        with tempfile.TemporaryDirectory() as tmp:
            for k, ops in enumerate(OPS):
                raw, (curves, points, slices) = run(meta, inp, ops, tmp)
                assert np.array_equal(raw["points"]["corner"], gold["points"]["corner"])     # the tracking fixture's punctured simplices
                trajs = raw["trajectories"]
                off = np.zeros(len(trajs) + 1, np.int64)
                for i, (idx, _) in enumerate(trajs):
                    off[i + 1] = off[i] + len(idx)
                arrays[f"trace_offsets_{k}"] = off
                arrays[f"trace_idx_{k}"] = np.concatenate([t[0] for t in trajs]).astype(np.int32) if trajs else np.zeros(0, np.int32)
                arrays[f"trace_loop_{k}"] = np.asarray([t[1] for t in trajs], np.uint8)
                # sliced critical points: (timestep, count) per slice + the concatenated point indices, in the reference's order
                arrays[f"slice_t_{k}"] = np.asarray([[t, len(ix)] for t, ix in slices], np.int64).reshape(-1, 2)
                arrays[f"slice_idx_{k}"] = np.concatenate([ix for _, ix in slices]).astype(np.int32) if slices else np.zeros(0, np.int32)
                arrays[f"curves_{k}"] = curves
                arrays[f"points_{k}"] = points[["idx", "cp_type", "ordinal", "timestep", "id", "t", "v"]]
                print(f"{name} [{ops}]: {len(curves)} curves, {len(points)} points")
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)


if __name__ == "__main__":
    main()
