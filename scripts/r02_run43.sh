#!/bin/bash
# round 2, GPU session 43 (1 GPU): same-box A/B of the 3D build kernel -- A: three unrolled plane steps, B: one step + register moves (a third of the
# code); ring depth / chunk length switches of the 2D key kernel and chunk length of the 3D kernel on the current library
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms/step %.4f scan %.4f frac %.3f value %.3e" % (d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["value"]), d.get("trajectories"), d.get("punctured_simplices"))
except Exception as e:
    print(sys.argv[2], "failed", e)
P
}
cp ftk_b200/libftkb200.so /tmp/lib_current.so
for v in A B A B; do
  cp ftk_b200/libvariant$v.so ftk_b200/libftkb200.so
  timeout 200 python bench.py --config c3 --only-main --steps 24 --warmup 4 --e2e-steps 0 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02t2_c3_$v.json
  show gpurun_out/r02t2_c3_$v.json "c3 variant $v"
done
cp ftk_b200/libvariantB.so ftk_b200/libftkb200.so
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "3d" 2>&1 | tail -2
cp /tmp/lib_current.so ftk_b200/libftkb200.so
for n in 4 5 6; do
  FTKB_K2_NST=$n timeout 200 python bench.py --config c2 --only-main --steps 60 --warmup 5 --e2e-steps 0 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02t2_c2_nst$n.json
  show gpurun_out/r02t2_c2_nst$n.json "c2 ring stages $n"
done
for r in 32 48; do
  FTKB_S3_ROWS=$r timeout 200 python bench.py --config c3 --only-main --steps 24 --warmup 4 --e2e-steps 0 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02t2_c3_rows$r.json
  show gpurun_out/r02t2_c3_rows$r.json "c3 planes per chunk $r"
done
