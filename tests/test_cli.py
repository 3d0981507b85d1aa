"""The C++ host layer: tracker classes (include/ftk_b200/critical_point_tracker_regular.hh) and the
`ftk -f cp`-compatible command line (ftk_b200/bin/ftkb200), both over the C ABI.

CPU: the header compiles as plain C++17 against ftkb200.h, the program links, prints its usage and fails
loudly without a GPU.  GPU: the program's outputs are compared with the reference fixtures / the oracle.
"""
import json
import os
import re
import subprocess

import numpy as np
import pytest

import _parity as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cli():
    from ftk_b200 import build
    build.build()
    assert os.path.exists(build.CLI)
    return build.CLI


def test_shim_header_is_self_contained(tmp_path):
    """a reference-style caller (python/pyftk.cpp:93-117 call order) compiles against the shim classes"""
    src = tmp_path / "caller.cpp"
    src.write_text(r'''
#include "ftk_b200/critical_point_tracker_regular.hh"
int run(const double *data, size_t DW, size_t DH, size_t DT) {
  ftk_b200::critical_point_tracker_2d_regular tracker;
  tracker.set_scalar_field_source(ftk_b200::SOURCE_GIVEN);
  tracker.set_vector_field_source(ftk_b200::SOURCE_DERIVED);
  tracker.set_jacobian_field_source(ftk_b200::SOURCE_DERIVED);
  tracker.set_jacobian_symmetric(true);
  tracker.set_domain(ftk_b200::lattice({2, 2}, {(int)DW - 3, (int)DH - 3}));
  tracker.set_array_domain(ftk_b200::lattice({0, 0}, {(int)DW, (int)DH}));
  tracker.initialize();
  for (size_t k = 0; k < DT; k++) {
    tracker.push_scalar_field_snapshot(ftk_b200::ndarray<double>::wrap(data + k * DW * DH, {DW, DH}));
    if (k != 0) tracker.advance_timestep();
    if (k == DT - 1) tracker.update_timestep();
  }
  tracker.finalize();
  int n = 0;
  for (const auto &kv : tracker.get_traced_critical_points()) n += (int)kv.second.size();
  return n + (int)tracker.get_critical_points().size();
}
''')
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)])


def test_cli_usage_and_loud_failure_without_gpu(cli, tmp_path):
    out = subprocess.run([cli, "--help"], capture_output=True, text=True)
    assert out.returncode == 0 and "--synthetic" in out.stdout and "--output-type" in out.stdout
    bad = subprocess.run([cli, "-f", "cp", "--synthetic", "woven"], capture_output=True, text=True)
    assert bad.returncode == 1 and "Missing '--output'" in bad.stderr
    bad = subprocess.run([cli, "-f", "cp", "--synthetic", "woven", "-o", str(tmp_path / "o.txt"), "-a", "none"], capture_output=True, text=True)
    assert bad.returncode == 1 and "no CPU path" in bad.stderr
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([cli, "-f", "cp", "--synthetic", "woven", "-o", str(tmp_path / "o.txt")], capture_output=True, text=True)
        assert r.returncode == 1 and "no CUDA device" in r.stderr and not os.path.exists(tmp_path / "o.txt")


# ---- GPU ------------------------------------------------------------------------------------------------
def _decode(points, dims, nv):
    """CLI JSON records -> parity record array; the 64-bit tag is (corner rank over domain x time) * ntypes + type"""
    nd = len(dims)
    lo = 2 if nv == 1 else 1
    size = [d - (3 if nv == 1 else 2) for d in dims]
    ntypes = 12 if nd == 2 else 60
    out = np.zeros(len(points), dtype=[("corner", np.int32, 4), ("simplex_type", np.int32), ("ordinal", np.int32), ("timestep", np.int32),
                                        ("cp_type", np.uint32), ("x", np.float64, 3), ("t", np.float64), ("scalar", np.float64)])
    for i, p in enumerate(points):
        tag = int(p["tag"])
        out["simplex_type"][i] = tag % ntypes
        r = tag // ntypes
        c = [0, 0, 0, 0]
        for j in range(nd):
            c[j] = r % size[j] + lo
            r //= size[j]
        c[3] = r
        out["corner"][i] = c
        out["ordinal"][i] = int(p["ordinal"])
        out["timestep"][i] = p["timestep"]
        out["cp_type"][i] = p["type"]
        out["x"][i] = [float("nan") if v is None else v for v in p["x"]]
        out["t"][i] = p["t"]
        out["scalar"][i] = float("nan") if p["scalar"][0] is None else p["scalar"][0]
    return out


def _sorted_like_reference(pts):
    order = np.lexsort((pts["simplex_type"], pts["corner"][:, 3], pts["corner"][:, 2], pts["corner"][:, 1], pts["corner"][:, 0]))
    return pts[order]


@pytest.mark.gpu
@pytest.mark.parametrize("name,args", [
    ("woven_cli_31x37x32", ["--synthetic", "woven", "--width", "31", "--height", "37", "--timesteps", "32"]),
    ("double_gyre_64x32x50", ["--synthetic", "double_gyre"]),
    ("merger_32x32x100", ["--synthetic", "merger_2d"]),
    ("mx3d_21x21x21x10", ["--synthetic", "moving_extremum_3d", "--timesteps", "10"]),
])
def test_cli_matches_reference_fixture(cli, tmp_path, name, args):
    meta, gold, _ = P.load_golden(name)
    if name.startswith("mx3d") and meta["params"]:
        args = args + ["--x0", ",".join(repr(float(v)) for v in meta["params"][:3]), "--dir", ",".join(repr(float(v)) for v in meta["params"][3:6])]
    dj, tt = tmp_path / "discrete.json", tmp_path / "traced.txt"
    r = subprocess.run([cli, "-f", "cp"] + args + ["-o", str(dj), "--output-type", "discrete", "--timing"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert re.search(r"t_init=.*t_compute=.*t_finalize=", r.stderr) and "kernel_launches=" in r.stderr
    got = _sorted_like_reference(_decode(json.load(open(dj)), meta["dims"], meta["nv"]))
    P.assert_same_result({"points": got, "trajectories": []}, gold, check_trajectories=False, tol=1e-9, what=f"cli {name}")
    r = subprocess.run([cli, "-f", "cp"] + args + ["-o", str(tt)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    text = open(tt).read().splitlines()
    assert text[0] == f"#trajectories={len(gold['trajectories'])}"
    lengths = sorted(len(idx) for idx, _ in gold["trajectories"])
    got_lengths, cur = [], None
    for line in text[1:]:
        if line.startswith("--trajectory"):
            if cur is not None:
                got_lengths.append(cur)
            cur = 0
        elif line.startswith("---"):
            cur += 1
    got_lengths.append(cur)
    assert sorted(got_lengths) == lengths


@pytest.mark.gpu
def test_cli_raw_input_matches_oracle(cli, tmp_path, oracle):
    """--input: one raw float64 file holding all timesteps of a 3D scalar series; traced JSON vs the oracle"""
    rng = np.random.default_rng(3)
    dims, T = [12, 10, 9], 4
    snaps = [rng.normal(size=(9, 10, 12)) for _ in range(T)]
    raw = tmp_path / "series.f64"
    np.stack(snaps).tofile(raw)
    o = oracle.track(snaps, dims, field="scalar")
    out = tmp_path / "traced.json"
    r = subprocess.run([cli, "-f", "cp", "--input", str(raw), "--input-format", "float64", "-w", "12", "-h", "10", "-d", "9", "-n", str(T), "-o", str(out)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    trajs = json.load(open(out))["trajs"]
    want_pts = o.points()
    want = sorted(tuple(tuple(int(v) for v in want_pts["corner"][i]) + (int(want_pts["simplex_type"][i]),) for i in idx) for idx, _ in o.trajectories())
    got = []
    for t in trajs:
        d = _decode(t["traj"], dims, 1)
        got.append(tuple(tuple(int(v) for v in d["corner"][i]) + (int(d["simplex_type"][i]),) for i in range(len(d))))
    assert sorted(got) == want
    # float32 input goes through the same path after widening
    raw32 = tmp_path / "series.f32"
    np.stack(snaps).astype(np.float32).tofile(raw32)
    o32 = oracle.track([s.astype(np.float32).astype(np.float64) for s in snaps], dims, field="scalar", trace=False)
    out32 = tmp_path / "d32.json"
    r = subprocess.run([cli, "-f", "cp", "--input", str(raw32), "--input-format", "float32", "-w", "12", "-h", "10", "-d", "9", "-n", str(T), "-o", str(out32),
                        "--output-type", "discrete"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got32 = _sorted_like_reference(_decode(json.load(open(out32)), dims, 1))
    P.assert_same_result({"points": got32, "trajectories": []}, {"points": o32.points(), "trajectories": []}, check_trajectories=False, tol=1e-9, what="cli float32")


@pytest.mark.gpu
def test_cli_post_process_matches_reference_curves(cli, tmp_path):
    """--post-process OPS (src/cli/ftk.cpp:253-257): merger_2d, traced text after the legacy op sequence; the curves
    (length, consistent type, loop flag) are those the unmodified reference produced for the same ops
    (tests/golden/post/merger_32x32x100.npz, op list 0)"""
    z = np.load(os.path.join(P.GOLDEN_DIR, "post", "merger_32x32x100.npz"))
    ops = json.loads(bytes(z["meta"]).decode())["ops"][0]
    out = tmp_path / "pp.txt"
    r = subprocess.run([cli, "-f", "cp", "--synthetic", "merger_2d", "--post-process", ops, "-o", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    text = open(out).read().splitlines()
    want = z["curves_0"]
    assert text[0] == f"#trajectories={len(want)}"
    got, cur = [], None
    for line in text[1:]:
        if line.startswith("--trajectory"):
            if cur is not None:
                got.append(tuple(cur))
            m = re.search(r"consistent_type=(\d+), loop=(\d+)", line)
            cur = [0, int(m.group(1)), int(m.group(2))]
        elif line.startswith("---"):
            cur[0] += 1
    if cur is not None:
        got.append(tuple(cur))
    assert sorted(got) == sorted((int(c["count"]), int(c["consistent_type"]), int(c["loop"])) for c in want)
    bad = subprocess.run([cli, "-f", "cp", "--synthetic", "merger_2d", "--post-process", "no_such_op", "-o", str(out)], capture_output=True, text=True)
    assert bad.returncode == 1 and "unknown operation" in bad.stderr


@pytest.mark.gpu
def test_cli_sliced_output(cli, tmp_path):
    """--output-type sliced (json_interface.hh:805-811): one text file per timestep holding the ordinal points of the traced
    curves; every punctured ordinal simplex of the reference fixture appears exactly once, in its own timestep's file"""
    meta, gold, _ = P.load_golden("merger_32x32x100")
    pat = str(tmp_path / "sliced-%03d.txt")
    r = subprocess.run([cli, "-f", "cp", "--synthetic", "merger_2d", "--output-type", "sliced", "-o", pat], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    pts = gold["points"]
    want = {int(t): int(n) for t, n in zip(*np.unique(pts["timestep"][pts["ordinal"] != 0], return_counts=True))}
    got = {}
    for t in range(meta["T"]):
        f = pat % t
        if os.path.exists(f):
            lines = open(f).read().splitlines()
            assert all(f"timestep={t}, ordinal=1" in line for line in lines)
            got[t] = len(lines)
    assert got == want


def test_shim_header_streaming_and_coords_compile(tmp_path):
    """the shim's streaming switch and the three coordinate setters (regular_tracker.hh:38-40, critical_point_tracker.hh:38)"""
    src = tmp_path / "caller2.cpp"
    src.write_text(r'''
#include "ftk_b200/critical_point_tracker_regular.hh"
int run(const double *data, size_t DW, size_t DH, size_t DD, const double *xs, const double *ys, const double *xy) {
  ftk_b200::critical_point_tracker_2d_regular t2;
  t2.set_enable_streaming_trajectories(true);
  t2.set_coords_bounds({0.0, 1.0, -1.0, 1.0});
  t2.set_coords_rectilinear({ftk_b200::ndarray<double>::wrap(xs, {DW}), ftk_b200::ndarray<double>::wrap(ys, {DH})});
  t2.set_coords_explicit(ftk_b200::ndarray<double>::wrap(xy, {2, DW, DH}));
  ftk_b200::critical_point_tracker_3d_regular t3;
  t3.set_coords_bounds({0.0, 1.0, 0.0, 1.0, 0.0, 1.0});
  t3.set_enable_streaming_trajectories(true);
  int n = 0;
  for (const auto &kv : t2.get_traced_critical_points()) n += kv.second.complete ? 1 : 0;
  return n + (int)(DD + (data != nullptr));
}
''')
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)])
