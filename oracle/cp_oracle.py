"""TEST INFRASTRUCTURE ONLY: ctypes wrapper of the plain-C parity oracle (oracle/cp_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
The product (ftk_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libcp_oracle.so")
REF_BINARY = os.path.join(_HERE, "_ref", "ftk_ref_oracle")

SOURCE_NONE, SOURCE_GIVEN, SOURCE_DERIVED = 0, 1, 2


def build(force=False):
    """Compile the C restatement (and, when /root/reference exists, the reference harness)."""
    src = os.path.join(_HERE, "cp_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "_build/libcp_oracle.so"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["bash", os.path.join(_HERE, "build_ref.sh")], stdout=subprocess.DEVNULL)


class Config(C.Structure):
    _fields_ = [
        ("nd", C.c_int32), ("dims", C.c_int32 * 3), ("lb", C.c_int32 * 3), ("ub", C.c_int32 * 3),
        ("scalar_source", C.c_int32), ("vector_source", C.c_int32), ("jacobian_source", C.c_int32),
        ("jacobian_symmetric", C.c_int32), ("robust_detection", C.c_int32), ("compute_degrees", C.c_int32),
        ("use_type_filter", C.c_int32), ("type_filter", C.c_uint32), ("start_timestep", C.c_int32),
        ("nthreads", C.c_int32),
    ]


POINT_DTYPE = np.dtype([
    ("corner", np.int32, 4), ("simplex_type", np.int32), ("ordinal", np.int32), ("timestep", np.int32),
    ("cp_type", np.uint32), ("x", np.float64, 3), ("t", np.float64), ("scalar", np.float64),
], align=True)
assert POINT_DTYPE.itemsize == 72

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.cpo_create.restype = C.c_void_p
        L.cpo_create.argtypes = [C.POINTER(Config)]
        for name in ("cpo_destroy",):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = None
        L.cpo_push_snapshot.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        for name in ("cpo_update_timestep", "cpo_advance_timestep", "cpo_finalize"):
            getattr(L, name).argtypes = [C.c_void_p]
        for name in ("cpo_scaling_factor", "cpo_resolution"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = C.c_double
        for name in ("cpo_num_points", "cpo_num_trajectories"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = C.c_uint64
        L.cpo_get_points.argtypes = [C.c_void_p, C.c_void_p]
        L.cpo_set_coords.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64]
        L.cpo_set_coords.restype = None
        L.cpo_set_resolution.argtypes = [C.c_void_p, C.c_double]
        L.cpo_set_resolution.restype = None
        L.cpo_import_points.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.cpo_import_points.restype = None
        L.cpo_get_trajectories.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.cpo_get_component_labels.argtypes = [C.c_void_p, C.c_void_p]
        L.cpo_get_degrees.argtypes = [C.c_void_p, C.c_void_p]
        L.cpo_array_resolution.restype = C.c_double
        L.cpo_array_resolution.argtypes = [C.c_void_p, C.c_uint64]
        L.cpo_gen_woven.argtypes = [C.c_int, C.c_int, C.c_double, C.c_void_p]
        L.cpo_gen_merger.argtypes = [C.c_int, C.c_int, C.c_double, C.c_void_p]
        L.cpo_gen_moving_extremum.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        L.cpo_gen_double_gyre.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p]
        L.cpo_gen_abc.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p]
        L.cpo_gen_tornado.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.cpo_cp_type_2d.restype = C.c_uint32
        L.cpo_cp_type_2d.argtypes = [C.c_void_p, C.c_int]
        L.cpo_cp_type_3d.restype = C.c_uint32
        L.cpo_cp_type_3d.argtypes = [C.c_void_p, C.c_int]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


# ---------------------------------------------------------------------------------------------
# generators.  Arrays are returned in *memory order* (dim 0 fastest), i.e. numpy shape is the
# reverse of the reference's ndarray shape: scalar -> ([D,] H, W), vector -> ([D,] H, W, n).
# ---------------------------------------------------------------------------------------------
def gen_woven(W, H, t):
    out = np.empty((H, W), np.float64)
    lib().cpo_gen_woven(W, H, float(t), _ptr(out))
    return out


def gen_merger(W, H, t):
    out = np.empty((H, W), np.float64)
    lib().cpo_gen_merger(W, H, float(t), _ptr(out))
    return out


def gen_moving_extremum(dims, x0, direction, t):
    nd = len(dims)
    d = np.asarray(dims, np.int32)
    out = np.empty(tuple(reversed(dims)), np.float64)
    lib().cpo_gen_moving_extremum(nd, _ptr(d), _ptr(np.asarray(x0, np.float64)), _ptr(np.asarray(direction, np.float64)),
                                  float(t), _ptr(out))
    return out


def gen_double_gyre(W, H, time, A=0.1, omega=2 * np.pi, eps=0.25):
    out = np.empty((H, W, 2), np.float64)
    lib().cpo_gen_double_gyre(W, H, float(time), A, omega, eps, _ptr(out))
    return out


def gen_abc(W, H, D, A=np.sqrt(3.0), B=np.sqrt(2.0), Cc=1.0):
    out = np.empty((D, H, W, 3), np.float64)
    lib().cpo_gen_abc(W, H, D, float(A), float(B), float(Cc), _ptr(out))
    return out


def abc_amplitude(k):
    """Unsteady-ABC modulation of SURVEY 8(d) C4 (ours): A(k) = sqrt(3) + 0.5 (k/64) sin(pi k/64)."""
    import math
    return math.sqrt(3.0) + 0.5 * (float(k) / 64.0) * math.sin(math.pi * float(k) / 64.0)


def gen_tornado(W, H, D, time):
    out = np.empty((D, H, W, 3), np.float64)
    lib().cpo_gen_tornado(W, H, D, int(time), _ptr(out))
    return out


def synthetic_series(name, dims, T, params=None):
    """Yield the T snapshots of a named reference generator (same time conventions as
    oracle/ref_harness.cpp)."""
    p = list(params or [])
    nd = len(dims)

    def P(i, d):
        return p[i] if i < len(p) else d
    for k in range(T):
        if name == "woven":
            yield gen_woven(dims[0], dims[1], (float(k) / (T - 1)) + 1e-4)
        elif name == "woven_cli":
            yield gen_woven(dims[0], dims[1], 0.0 if T == 1 else float(k) / (T - 1))
        elif name == "merger":
            yield gen_merger(dims[0], dims[1], float(k) * 0.1)
        elif name == "moving_extremum" and nd == 2:
            yield gen_moving_extremum(dims, [P(0, 10), P(1, 10)], [P(2, 0.1), P(3, 0.1)], float(k))
        elif name == "moving_extremum" and nd == 3:
            yield gen_moving_extremum(dims, [P(0, 10), P(1, 10), P(2, 10)], [P(3, 0.1), P(4, 0.11), P(5, 0.1)], float(k))
        elif name == "double_gyre":
            yield gen_double_gyre(dims[0], dims[1], k * P(0, 0.1))
        elif name == "abc":
            yield gen_abc(dims[0], dims[1], dims[2], abc_amplitude(k))
        elif name == "tornado":
            yield gen_tornado(dims[0], dims[1], dims[2], k)
        else:
            raise ValueError(name)


# ---------------------------------------------------------------------------------------------
COORDS_MODES = {"bounds": 1, "rectilinear": 2, "explicit": 3}


class Tracker:
    """Oracle twin of ftk::critical_point_tracker_{2d,3d}_regular (CPU, non-GMP)."""

    def __init__(self, dims, lb=None, ub=None, field="scalar", jacobian_symmetric=None, robust=True,
                 compute_degrees=False, type_filter=None, start_timestep=0, nthreads=0,
                 scalar_source=None, vector_source=None, jacobian_source=None, coords=None):
        nd = len(dims)
        cfg = Config()
        cfg.nd = nd
        for i in range(nd):
            cfg.dims[i] = dims[i]
            # defaults of json_interface.hh:639-654
            cfg.lb[i] = (2 if field == "scalar" else 1) if lb is None else lb[i]
            cfg.ub[i] = dims[i] - 2 if ub is None else ub[i]
        if nd == 2:
            cfg.dims[2] = 1
        if field == "scalar":
            ss, vs, js = SOURCE_GIVEN, SOURCE_DERIVED, SOURCE_DERIVED
            sym = True if jacobian_symmetric is None else jacobian_symmetric
        else:
            ss, vs, js = SOURCE_NONE, SOURCE_GIVEN, SOURCE_DERIVED
            sym = False if jacobian_symmetric is None else jacobian_symmetric
        cfg.scalar_source = ss if scalar_source is None else scalar_source
        cfg.vector_source = vs if vector_source is None else vector_source
        cfg.jacobian_source = js if jacobian_source is None else jacobian_source
        cfg.jacobian_symmetric = int(sym)
        cfg.robust_detection = int(robust)
        cfg.compute_degrees = int(compute_degrees)
        cfg.use_type_filter = int(type_filter is not None)
        cfg.type_filter = int(type_filter or 0)
        cfg.start_timestep = start_timestep
        cfg.nthreads = nthreads
        self.cfg = cfg
        self.nd = nd
        self._h = lib().cpo_create(C.byref(cfg))
        if not self._h:
            raise RuntimeError("cpo_create failed")
        if coords is not None:      # (mode name, flat float64 data): regular_tracker.hh:38-40
            mode, data = coords
            data = np.ascontiguousarray(data, np.float64).ravel()
            lib().cpo_set_coords(self._h, COORDS_MODES[mode], _ptr(data), data.size)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().cpo_destroy(self._h)
            self._h = None

    def push_field_data_snapshot(self, scalar=None, vector=None, jacobian=None):
        s, v, j = _f64(scalar), _f64(vector), _f64(jacobian)
        rc = lib().cpo_push_snapshot(self._h, _ptr(s), _ptr(v), _ptr(j))
        if rc:
            raise RuntimeError("cpo_push_snapshot failed")

    def push_scalar_field_snapshot(self, s):
        self.push_field_data_snapshot(scalar=s)

    def push_vector_field_snapshot(self, v):
        self.push_field_data_snapshot(vector=v)

    def update_timestep(self):
        lib().cpo_update_timestep(self._h)

    def advance_timestep(self):
        lib().cpo_advance_timestep(self._h)

    def finalize(self):
        lib().cpo_finalize(self._h)

    @property
    def scaling_factor(self):
        return lib().cpo_scaling_factor(self._h)

    @property
    def resolution(self):
        return lib().cpo_resolution(self._h)

    def points(self):
        n = lib().cpo_num_points(self._h)
        out = np.zeros(n, POINT_DTYPE)
        if n:
            lib().cpo_get_points(self._h, _ptr(out))
        return out

    def trajectories(self):
        """-> list of (index array into points(), loop flag)"""
        nt = lib().cpo_num_trajectories(self._h)
        npts = lib().cpo_num_points(self._h)
        off = np.zeros(nt + 1, np.uint64)
        idx = np.zeros(max(npts, 1), np.uint64)
        loop = np.zeros(max(nt, 1), np.uint8)
        if nt:
            lib().cpo_get_trajectories(self._h, _ptr(off), _ptr(idx), _ptr(loop))
        return [(idx[int(off[i]):int(off[i + 1])].astype(np.int64), bool(loop[i])) for i in range(nt)]

    # time-slab helpers (multi-GPU protocol tests)
    def set_resolution(self, r):
        lib().cpo_set_resolution(self._h, float(r))

    def import_points(self, pts):
        pts = np.ascontiguousarray(pts, dtype=POINT_DTYPE)
        if len(pts):
            lib().cpo_import_points(self._h, _ptr(pts), len(pts))

    def component_labels(self):
        n = lib().cpo_num_points(self._h)
        out = np.zeros(max(n, 1), np.uint64)
        lib().cpo_get_component_labels(self._h, _ptr(out))
        return out[:n]

    def degrees(self):
        n = lib().cpo_num_points(self._h)
        out = np.zeros(max(n, 1), np.int32)
        lib().cpo_get_degrees(self._h, _ptr(out))
        return out[:n]


def track(snapshots, dims, field="scalar", trace=True, **kw):
    """Run the reference's front-end loop (python/pyftk.cpp:110-117) over an iterable of snapshots."""
    tr = Tracker(dims, field=field, **kw)
    snaps = list(snapshots)
    T = len(snaps)
    for k, s in enumerate(snaps):
        if field == "scalar":
            tr.push_scalar_field_snapshot(s)
        else:
            tr.push_vector_field_snapshot(s)
        if k != 0:
            tr.advance_timestep()
        if k == T - 1:
            tr.update_timestep()
    if trace:
        tr.finalize()
    return tr


# ---------------------------------------------------------------------------------------------
# golden fixtures written by oracle/ref_harness.cpp (.ftkg)
# ---------------------------------------------------------------------------------------------
def read_ftkg(path):
    import gzip
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rb") as f:
        buf = f.read()
    magic, version, nd, traced = np.frombuffer(buf, np.uint32, 4, 0)
    assert magic == 0x474B5446 and version == 1
    npts, ntraj = (int(v) for v in np.frombuffer(buf, np.uint64, 2, 16))
    rec = np.dtype([("corner", np.int32, 4), ("simplex_type", np.int32), ("ordinal", np.int32),
                    ("timestep", np.int32), ("cp_type", np.int32), ("x", np.float64, 3), ("t", np.float64),
                    ("scalar", np.float64)])
    assert rec.itemsize == 72
    pts = np.frombuffer(buf, rec, npts, 32)
    off = 32 + 72 * npts
    trajs = []
    for _ in range(ntraj):
        ln, loop = (int(v) for v in np.frombuffer(buf, np.uint64, 2, off))
        off += 16
        idx = np.frombuffer(buf, np.uint64, ln, off).astype(np.int64)
        off += 8 * ln
        trajs.append((idx, bool(loop)))
    return {"nd": int(nd), "traced": bool(traced), "points": pts, "trajectories": trajs}


def canonical_trajectories(trajs):
    """Order-independent form: set of (tuple(indices), loop).  A trajectory and its reverse are
    different sequences; the reference's walk is deterministic given the component, so sequences
    are compared as-is."""
    return sorted((tuple(int(i) for i in idx), bool(loop)) for idx, loop in trajs)


def run_reference(nd, nv, dims, T, gen=None, params=None, input_array=None, out=None, trace=True, domain=None,
                  symmetric=None, nthreads=0, timeout=3600, coords=None, start_timestep=0):
    """Run the unmodified reference binary (oracle/_ref).  Returns (stats dict, golden dict or None)."""
    import json
    import tempfile
    if not os.path.exists(REF_BINARY):
        raise FileNotFoundError(REF_BINARY)
    cmd = [REF_BINARY, "--nd", str(nd), "--nv", str(nv), "--dims"] + [str(d) for d in dims] + ["--nt", str(T), "--quiet"]
    tmp_in = None
    if input_array is not None:
        tmp_in = tempfile.NamedTemporaryFile(suffix=".f64", delete=False)
        np.ascontiguousarray(input_array, np.float64).tofile(tmp_in)
        tmp_in.close()
        cmd += ["--input", tmp_in.name]
    else:
        cmd += ["--gen", gen]
        if params:
            cmd += ["--p"] + [repr(float(x)) for x in params]
    if domain is not None:
        cmd += ["--domain"] + [str(int(v)) for v in domain]
    if symmetric is not None:
        cmd += ["--symmetric", str(int(symmetric))]
    if nthreads:
        cmd += ["--nthreads", str(nthreads)]
    if start_timestep:
        cmd += ["--start-timestep", str(int(start_timestep))]
    if not trace:
        cmd += ["--no-trace"]
    tmp_coords = None
    if coords is not None:
        tmp_coords = tempfile.NamedTemporaryFile(suffix=".f64", delete=False)
        np.ascontiguousarray(coords[1], np.float64).ravel().tofile(tmp_coords)
        tmp_coords.close()
        cmd += ["--coords", coords[0], "--coords-file", tmp_coords.name]
    tmp_out = None
    if out is None:
        tmp_out = tempfile.NamedTemporaryFile(suffix=".ftkg", delete=False)
        tmp_out.close()
        out = tmp_out.name
    cmd += ["--out", out]
    try:
        res = subprocess.run(cmd, check=True, capture_output=True, timeout=timeout)
        stats = json.loads(res.stdout.decode().strip().splitlines()[-1])
        gold = read_ftkg(out)
    finally:
        if tmp_in is not None:
            os.unlink(tmp_in.name)
        if tmp_coords is not None:
            os.unlink(tmp_coords.name)
        if tmp_out is not None and os.path.exists(tmp_out.name):
            os.unlink(tmp_out.name)
    return stats, gold
