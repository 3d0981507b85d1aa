"""CPU: host-side logic of the product (mesh tables, C-ABI surface, input marshalling) -- no GPU compute."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ftkb():
    from ftk_b200 import build
    build.build()
    from ftk_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol(ftkb):
    """every function include/ftkb200.h declares is exported by libftkb200.so (and listed in the binding)"""
    header = open(os.path.join(ROOT, "include", "ftkb200.h")).read()
    declared = sorted(set(re.findall(r"\b(ftkb_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations found"
    lib = ftkb.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in ftkb200.h but not exported"
    assert sorted(ftkb.EXPORTS) == declared
    assert lib.ftkb_abi_version() == ftkb.ABI_VERSION


def test_struct_layouts_match_header(ftkb, tmp_path):
    """the ctypes mirrors have the sizes a C compiler gives the structs of include/ftkb200.h (the header is plain C)"""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "ftkb200.h"\nint main(void){printf("%zu %zu %zu\\n", sizeof(ftkb_config), sizeof(ftkb_point), sizeof(ftkb_stats));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(ftkb.Config), ftkb.POINT_DTYPE.itemsize, C.sizeof(ftkb.Stats)]
    assert ftkb.POINT_DTYPE.itemsize == 72


def test_no_cpu_fallback(ftkb):
    """without a usable device ftkb_create fails with FTKB_ERR_NO_DEVICE instead of computing on the CPU"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cfg = ftkb.Config()
    cfg.abi_version, cfg.nd = ftkb.ABI_VERSION, 2
    for i, (d, lo, hi) in enumerate([(16, 2, 14), (16, 2, 14), (1, 0, 0)]):
        cfg.dims[i], cfg.lb[i], cfg.ub[i] = d, lo, hi
    cfg.scalar_source, cfg.vector_source, cfg.jacobian_source = 1, 2, 2
    h = C.c_void_p()
    assert ftkb.lib().ftkb_create(C.byref(cfg), C.byref(h)) == ftkb.ERR_NO_DEVICE
    assert b"no CPU fallback" in ftkb.lib().ftkb_last_error(None) or b"CUDA device" in ftkb.lib().ftkb_last_error(None)
    assert ftkb.lib().ftkb_device_count() == 0


def test_create_rejects_bad_arguments(ftkb):
    cfg = ftkb.Config()
    cfg.abi_version, cfg.nd = ftkb.ABI_VERSION + 1, 2
    h = C.c_void_p()
    assert ftkb.lib().ftkb_create(C.byref(cfg), C.byref(h)) == ftkb.ERR_INVALID
    cfg.abi_version, cfg.nd = ftkb.ABI_VERSION, 4
    assert ftkb.lib().ftkb_create(C.byref(cfg), C.byref(h)) == ftkb.ERR_INVALID
    cfg.nd = 2
    cfg.dims[0], cfg.dims[1], cfg.lb[0], cfg.ub[0], cfg.lb[1], cfg.ub[1] = 8, 8, 2, 9, 2, 6   # ub outside the array
    assert ftkb.lib().ftkb_create(C.byref(cfg), C.byref(h)) == ftkb.ERR_INVALID


def _table(fn, nd, k, t, cap=128):
    buf = (C.c_int32 * (cap * (nd + 1)))()
    n = fn(nd, k, t, buf)
    return sorted(tuple(buf[i * (nd + 1) + j] for j in range(nd + 1)) for i in range(n))


@pytest.mark.parametrize("nd", [3, 4])
def test_mesh_tables_match_oracle(ftkb, oracle, nd):
    """unit simplices, ordinal/interval split, facets and cofaces == the oracle's restatement of
    simplicial_regular_mesh.hh:620-831 (itself pinned on the reference through the golden fixtures)"""
    P, O = ftkb.lib(), oracle.lib()
    for k in range(nd + 1):
        nt = P.ftkb_mesh_ntypes(nd, k, 0)
        assert nt == O.cpo_mesh_ntypes(nd, k, 0)
        for scope in (1, 2):
            ns = P.ftkb_mesh_ntypes(nd, k, scope)
            assert ns == O.cpo_mesh_ntypes(nd, k, scope)
            assert [P.ftkb_mesh_scope_type(nd, k, scope, i) for i in range(ns)] == \
                   [O.cpo_mesh_scope_type(nd, k, scope, i) for i in range(ns)]
        for t in range(nt):
            a, b = (C.c_int32 * ((k + 1) * nd))(), (C.c_int32 * ((k + 1) * nd))()
            assert P.ftkb_mesh_unit_simplex(nd, k, t, a) == 0
            O.cpo_mesh_unit_simplex(nd, k, t, b)
            assert list(a) == list(b), (nd, k, t)
            assert _table(P.ftkb_mesh_sides, nd, k, t) == _table(O.cpo_mesh_sides, nd, k, t)
            assert _table(P.ftkb_mesh_side_of, nd, k, t) == _table(O.cpo_mesh_side_of, nd, k, t)


def test_mesh_table_counts(ftkb):
    """SURVEY.md App. C"""
    P = ftkb.lib()
    assert [P.ftkb_mesh_ntypes(3, k, 0) for k in range(4)] == [1, 7, 12, 6]
    assert [P.ftkb_mesh_ntypes(4, k, 0) for k in range(5)] == [1, 15, 50, 60, 24]
    assert [P.ftkb_mesh_scope_type(3, 2, 1, i) for i in range(2)] == [4, 8]
    assert [P.ftkb_mesh_scope_type(4, 3, 1, i) for i in range(6)] == [16, 20, 30, 34, 46, 50]
    # every n-simplex of the (n+1)-D mesh has exactly two cofaces; every cell n+2 facets
    for nd in (3, 4):
        for t in range(P.ftkb_mesh_ntypes(nd, nd - 1, 0)):
            assert len(_table(P.ftkb_mesh_side_of, nd, nd - 1, t)) == 2
        for t in range(P.ftkb_mesh_ntypes(nd, nd, 0)):
            assert len(_table(P.ftkb_mesh_sides, nd, nd, t)) == nd + 1


def test_sign_early_out_is_exact(oracle):
    """The scan kernel's early-out: if every vertex of a simplex is strictly on one side of zero in
    some component, the reference predicate (oracle restatement of sign_det.hh) returns false.
    Degenerate-heavy random integer simplices with distinct random ranks."""
    O = oracle.lib()
    O.cpo_robust_cp_in_simplex2.argtypes = [C.c_void_p, C.c_void_p]
    O.cpo_robust_cp_in_simplex3.argtypes = [C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(7)
    for nv, nc, fn in ((3, 2, O.cpo_robust_cp_in_simplex2), (4, 3, O.cpo_robust_cp_in_simplex3)):
        punctured = sided = 0
        for r in (1, 2, 3, 4):
            X = rng.integers(-r, r + 1, size=(20000, nv, nc)).astype(np.int64)
            idx = np.argsort(rng.random((20000, nv)), axis=1).astype(np.int32) + rng.integers(0, 1000, size=(20000, 1)).astype(np.int32)
            one_sided = ((X > 0).all(axis=1) | (X < 0).all(axis=1)).any(axis=1)
            for i in range(len(X)):
                hit = fn(X[i].ctypes.data, idx[i].ctypes.data)
                punctured += hit
                if one_sided[i]:
                    sided += 1
                    assert not hit, (X[i], idx[i])
        assert punctured > 1000 and sided > 1000


def test_pyftk_input_marshalling():
    """track_critical_points_2d_scalar reads the numpy buffer dim-0-fastest (python/pyftk.cpp:97-99)"""
    from ftk_b200 import synthesizers as S
    a = S.spiral_woven(6, 5, 3)
    assert a.shape == (1, 6, 5, 3)
    snap1 = a.reshape(-1)[30:60].reshape(5, 6)
    assert np.array_equal(snap1, S.woven_snapshot(6, 5, 1.0 / 2 + 1e-4))
    assert S.moving_extremum(7, 6, 4, 3, 3, 0.1, 0.2).shape == (1, 7, 6, 4)
    assert S.double_gyre_flow(8, 4, 2).shape == (2, 8, 4, 2)


def test_python_generators_match_oracle(oracle):
    from ftk_b200 import synthesizers as S
    assert np.allclose(S.woven_snapshot(9, 7, 0.3), oracle.gen_woven(9, 7, 0.3), rtol=0, atol=1e-15)
    assert np.allclose(S.double_gyre_snapshot(9, 7, 0.3), oracle.gen_double_gyre(9, 7, 0.3), rtol=0, atol=1e-15)
    assert np.array_equal(S.moving_extremum_snapshot([9, 7], [4, 3], [0.1, 0.2], 2.0), oracle.gen_moving_extremum([9, 7], [4, 3], [0.1, 0.2], 2.0))
    assert np.array_equal(S.moving_extremum_snapshot([6, 5, 4], [3, 2, 1], [0.1, 0.2, 0.3], 2.0),
                          oracle.gen_moving_extremum([6, 5, 4], [3, 2, 1], [0.1, 0.2, 0.3], 2.0))
