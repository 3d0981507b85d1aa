"""Generate tests/golden/binary/*.npz: the reference's own binary archives (DIY serialization) of discrete and traced
critical points, written by the UNMODIFIED reference (oracle/_ref/ftk_ref_oracle --binary-discrete / --binary-traced,
i.e. critical_point_tracker::write_critical_points_binary / write_traced_critical_points_binary).

    python tests/golden/make_golden_binary.py       (build container: needs /root/reference for oracle/build_ref.sh)
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
OUT = os.path.join(HERE, "binary")
CASES = ["mx3d_21x21x21x10", "mx2d_11x13x20", "rand2d_scalar_int", "double_gyre_64x32x50"]


def main():
    O.build()
    os.makedirs(OUT, exist_ok=True)
    for name in CASES:
        meta, gold, inp = P.load_golden(name)
        cmd = [O.REF_BINARY, "--nd", str(meta["nd"]), "--nv", str(meta["nv"]), "--dims"] + [str(d) for d in meta["dims"]] + ["--nt", str(meta["T"]), "--quiet"]
        with tempfile.TemporaryDirectory() as tmp:
            if inp is not None:
                raw = os.path.join(tmp, "in.f64")
                np.ascontiguousarray(inp, np.float64).tofile(raw)
                cmd += ["--input", raw]
            else:
                cmd += ["--gen", meta["gen"]]
                if meta["params"]:
                    cmd += ["--p"] + [repr(float(x)) for x in meta["params"]]
            d, t = os.path.join(tmp, "d.bin"), os.path.join(tmp, "t.bin")
            subprocess.run(cmd + ["--out", os.path.join(tmp, "o.ftkg"), "--binary-discrete", d, "--binary-traced", t], check=True, capture_output=True)
            db, tb = open(d, "rb").read(), open(t, "rb").read()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=np.frombuffer(json.dumps(dict(case=name)).encode(), np.uint8),
                            discrete=np.frombuffer(db, np.uint8), traced=np.frombuffer(tb, np.uint8))
        print(f"{name}: discrete {len(db)} B, traced {len(tb)} B")


if __name__ == "__main__":
    main()
