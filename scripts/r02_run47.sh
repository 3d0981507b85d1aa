#!/bin/bash
# round 2, GPU session 47 (1 GPU): same-box A/B of the 3D kernels -- B keeps the unused fourth word of the other layer's cell alive until the cell
# is consumed (A: the compiler reuses it while the 128-bit load is in flight)
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
  timeout 200 python bench.py --config c3 --only-main --steps 24 --warmup 4 --e2e-steps 0 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02w2_c3_$v.json
  show gpurun_out/r02w2_c3_$v.json "c3 variant $v"
done
for v in A B; do
  cp ftk_b200/libvariant$v.so ftk_b200/libftkb200.so
  timeout 400 python bench.py --config c4 --only-main --steps 8 --warmup 3 --e2e-steps 0 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02w2_c4_$v.json
  show gpurun_out/r02w2_c4_$v.json "c4 variant $v"
done
cp /tmp/lib_current.so ftk_b200/libftkb200.so
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_group.py -m gpu -q -x -k "3d or slab or abc or tornado" 2>&1 | tail -2
