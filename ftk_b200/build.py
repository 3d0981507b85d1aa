"""Build the C-ABI shared library (libftkb200.so) in-tree with nvcc for sm_100a.

    python -m ftk_b200.build [--force]

The library has no Python or torch dependency; ftk_b200/_lib.py loads it with ctypes.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libftkb200.so")
CLI = os.path.join(HERE, "bin", "ftkb200")          # `ftk -f cp` front end (C++ host code over the C ABI)
CLI_SOURCES = [os.path.join(CSRC, "cli.cpp"), os.path.join(HERE, "..", "include", "ftk_b200", "critical_point_tracker_regular.hh"),
               os.path.join(HERE, "..", "include", "ftkb200.h")]
SOURCES = ["kernels.cu", "context.cpp", "mesh_tables.cpp", "curves.cpp", "online.cpp", "group.cpp"]
HEADERS = ["kernels.h", "mesh_tables.h", "curves.h", "online.h", os.path.join("..", "..", "include", "ftkb200.h")]
# -fmad=false: the parity target is the reference's x86-64 CPU path (no FMA contraction)
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
              "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function", "-shared", "--cudart", "static", "-Xcompiler", "-pthread"]


def nvcc():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    raise RuntimeError("nvcc not found")


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS) or os.path.getmtime(__file__) > t


def build_cli(force=False):
    """g++ host program linked against the in-tree library (rpath $ORIGIN/..)"""
    if not force and os.path.exists(CLI) and all(os.path.getmtime(f) <= os.path.getmtime(CLI) for f in CLI_SOURCES + [LIB]):
        return CLI
    os.makedirs(os.path.dirname(CLI), exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-pthread", CLI_SOURCES[0], "-o", CLI, "-L" + HERE, "-lftkb200", "-Wl,-rpath,$ORIGIN/.."]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed: " + " ".join(cmd))
    return CLI


def build(force=False, verbose=False):
    _build_lib(force, verbose)
    build_cli(force)
    return LIB


def _build_lib(force=False, verbose=False):
    if not force and not stale():
        return LIB
    cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, f) for f in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
