"""Binary archives: the reference's DIY serialization of discrete and traced critical points
(critical_point_tracker.hh:339-364, feature_point.hh:160-188, feature_curve.hh:472-507, feature_curve_set.hh:92-112).

CPU: archives written by the UNMODIFIED reference (tests/golden/binary/*.npz, made by tests/golden/make_golden_binary.py)
are read by the C++ shim classes and written back: the bytes must be identical; an independent Python parser of the
format agrees with the tracking fixtures.  GPU: the CLI's --output-format binary, parsed, equals the fixtures (tags
included: they are the reference's element ids).
"""
import glob
import os
import struct
import subprocess

import numpy as np
import pytest

import _parity as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN_DIR = os.path.join(P.GOLDEN_DIR, "binary")
NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(BIN_DIR, "*.npz")))
POINT = struct.Struct("<3d d i 3d 3d I B Q Q")      # 105 bytes, no padding
assert POINT.size == 105


def parse_points(buf, off, n):
    out = []
    for _ in range(n):
        f = POINT.unpack_from(buf, off)
        off += POINT.size
        out.append(dict(x=f[0:3], t=f[3], timestep=f[4], scalar=f[5:8], v=f[8:11], type=f[11], ordinal=f[12], tag=f[13], id=f[14]))
    return out, off


def parse_discrete(buf):
    n = struct.unpack_from("<Q", buf, 0)[0]
    pts, off = parse_points(buf, 8, n)
    assert off == len(buf)
    return pts


def parse_traced(buf):
    n = struct.unpack_from("<Q", buf, 0)[0]
    off, curves = 8, []
    for _ in range(n):
        cid, complete = struct.unpack_from("<iB", buf, off)
        off += 5
        stats = struct.unpack_from("<17d", buf, off)
        off += 136
        ctype, npts = struct.unpack_from("<IQ", buf, off)
        off += 12
        pts, off = parse_points(buf, off, npts)
        curves.append(dict(id=cid, complete=complete, stats=stats, consistent_type=ctype, points=pts))
    assert off == len(buf)
    return curves


@pytest.fixture(scope="module")
def roundtrip_tool(tmp_path_factory):
    """a reference-style caller of the shim classes: read an archive, write it back"""
    d = tmp_path_factory.mktemp("bin")
    src = d / "rt.cpp"
    src.write_text(r'''
#include "ftk_b200/critical_point_tracker_regular.hh"
int main(int argc, char **argv) {
  ftk_b200::critical_point_tracker_2d_regular tracker;       // no initialize(): archives need no device
  const std::string kind = argv[1];
  if (kind == "discrete") { if (!tracker.read_critical_points_binary(argv[2])) return 2; tracker.write_critical_points_binary(argv[3]); }
  else { if (!tracker.read_traced_critical_points_binary(argv[2])) return 2; tracker.write_traced_critical_points_binary(argv[3]); }
  return 0;
}
''')
    exe = d / "rt"
    lib = os.path.join(ROOT, "ftk_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L" + lib, "-lftkb200", "-Wl,-rpath," + lib])
    return str(exe)


@pytest.mark.parametrize("name", NAMES)
def test_reference_archives_round_trip_through_the_shim(name, roundtrip_tool, tmp_path):
    z = np.load(os.path.join(BIN_DIR, name + ".npz"))
    for kind in ("discrete", "traced"):
        src, dst = tmp_path / f"{kind}.in", tmp_path / f"{kind}.out"
        src.write_bytes(bytes(z[kind]))
        subprocess.check_call([roundtrip_tool, kind, str(src), str(dst)])
        assert dst.read_bytes() == bytes(z[kind]), f"{name}: {kind} archive changed in a read / write round trip"


@pytest.mark.parametrize("name", NAMES)
def test_reference_archives_hold_the_fixture_points(name):
    """the format as parsed here is the format the reference wrote: same points as the tracking fixture, same curves"""
    z = np.load(os.path.join(BIN_DIR, name + ".npz"))
    meta, gold, _ = P.load_golden(name)
    pts = parse_discrete(bytes(z["discrete"]))
    g = gold["points"]
    assert len(pts) == len(g)
    assert [p["timestep"] for p in pts] == [int(v) for v in g["timestep"]] and [p["type"] for p in pts] == [int(v) for v in g["cp_type"]]
    assert np.array_equal(np.asarray([p["x"] for p in pts]), g["x"]) and np.array_equal(np.asarray([p["t"] for p in pts]), g["t"])
    assert len(set(p["tag"] for p in pts)) == len(pts)
    curves = parse_traced(bytes(z["traced"]))
    assert len(curves) == len(gold["trajectories"]) and sorted(len(c["points"]) for c in curves) == sorted(len(i) for i, _ in gold["trajectories"])
    assert [c["id"] for c in curves] == list(range(len(curves))) and all(q["id"] == c["id"] for c in curves for q in c["points"])


@pytest.mark.gpu
@pytest.mark.parametrize("name,args", [
    ("double_gyre_64x32x50", ["--synthetic", "double_gyre"]),
    ("mx3d_21x21x21x10", ["--synthetic", "moving_extremum_3d", "--timesteps", "10"]),
])
def test_cli_binary_output_matches_reference_archive(name, args, tmp_path):
    from ftk_b200 import build
    build.build()
    z = np.load(os.path.join(BIN_DIR, name + ".npz"))
    want = parse_discrete(bytes(z["discrete"]))
    out = tmp_path / "d.bin"
    r = subprocess.run([build.CLI, "-f", "cp"] + args + ["--output-type", "discrete", "--output-format", "binary", "-o", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = parse_discrete(out.read_bytes())
    assert len(got) == len(want)
    key = lambda p: p["tag"]
    for a, b in zip(sorted(got, key=key), sorted(want, key=key)):
        assert (a["tag"], a["timestep"], a["type"], a["ordinal"]) == (b["tag"], b["timestep"], b["type"], b["ordinal"])
        assert max(abs(u - v) for u, v in zip(a["x"] + (a["t"], a["scalar"][0]), b["x"] + (b["t"], b["scalar"][0]))) <= 1e-9
    tout = tmp_path / "t.bin"
    r = subprocess.run([build.CLI, "-f", "cp"] + args + ["--output-format", "binary", "-o", str(tout)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    gc, wc = parse_traced(tout.read_bytes()), parse_traced(bytes(z["traced"]))
    assert sorted(tuple(q["tag"] for q in c["points"]) for c in gc) == sorted(tuple(q["tag"] for q in c["points"]) for c in wc)
    assert all(c["complete"] == 0 for c in gc)
