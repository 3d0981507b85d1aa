"""origin_in_simplex_fast (csrc/kernels.cu) against the full sort + symbolic-perturbation cascade it short-cuts.

The exact predicates are plain integer C++ (mod-2^64 arithmetic), so the section between the `[host-testable]` markers is
compiled for the host with g++ and both forms are run on random inputs: tiny values (exact ties in every determinant), values
whose products wrap int64, powers of two (determinants of exactly -2^63), duplicated rows.  The oracle comparison of the whole
tracker on the GPU is in test_gpu_parity.py; this test pins the identity the fast path rests on (the determinant is an
alternating polynomial, so sorting rows by rank and correcting by the swap parity gives the sign of the unsorted determinant
unless that determinant is 0 or -2^63) without a device."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = r'''
#include <cstdint>
#include <cstdio>
#include <climits>
#include <random>
#define __device__
#define __forceinline__ inline
#define __noinline__
typedef long long i64;
typedef unsigned long long u64;
#include "section.inc"
int main() {
  std::mt19937_64 rng(12345);
  long bad = 0, inside2 = 0, inside3 = 0, n = 0;
  auto val = [&](int mode) -> i64 {
    switch (mode) {
      case 0: return (i64)(rng() % 7) - 3;
      case 1: return (i64)(rng() % 2001) - 1000;
      case 2: return (i64)(rng() >> 20) - (1ll << 43);
      case 3: return (i64)rng();
      default: return (rng() & 1) ? (i64)(1ull << 63) : (i64)((rng() % 5) << 62);
    }
  };
  for (int it = 0; it < 1500000; it++) {
    const int mode = it % 5;
    i64 X[3][2]; int idx[3];
    for (int i = 0; i < 3; i++) { X[i][0] = val(mode); X[i][1] = val(mode); idx[i] = (int)(rng() % 1000); }
    if (it % 7 == 0) { X[1][0] = X[0][0]; X[1][1] = X[0][1]; }
    const bool a = origin_in_simplex<3, 2>(X, idx), b = origin_in_simplex_fast(X, idx);
    bad += a != b; inside2 += a;
    i64 Y[4][3]; int id4[4];
    for (int i = 0; i < 4; i++) { for (int j = 0; j < 3; j++) Y[i][j] = val(mode); id4[i] = (int)(rng() % 1000); }
    if (it % 11 == 0) for (int j = 0; j < 3; j++) Y[2][j] = Y[0][j];
    const bool c = origin_in_simplex<4, 3>(Y, id4), d = origin_in_simplex_fast(Y, id4);
    bad += c != d; inside3 += c;
    n++;
  }
  printf("%ld cases each, inside 2D %ld, inside 3D %ld, mismatches %ld\n", n, inside2, inside3, bad);
  return bad != 0 || inside2 == 0 || inside3 == 0;
}
'''


def test_fast_origin_in_simplex_equals_the_full_cascade(tmp_path):
    src = open(os.path.join(ROOT, "ftk_b200", "csrc", "kernels.cu")).read()
    a, b = src.index("// [host-testable: begin]"), src.index("// [host-testable: end]")
    (tmp_path / "section.inc").write_text(src[a:b])
    (tmp_path / "harness.cpp").write_text(HARNESS)
    exe = tmp_path / "harness"
    res = subprocess.run(["g++", "-O2", "-std=c++17", "-I", str(tmp_path), str(tmp_path / "harness.cpp"), "-o", str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "mismatches 0" in run.stdout
