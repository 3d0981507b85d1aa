#!/bin/bash
# round 2, GPU session 42 (1 GPU): same-box A/B of the 2D key kernel -- A: the other layer's cell loaded where it is used, B: staged with cp.async;
# B2: B + the successor CTA's first two stages prefetched into L2 (FTKB_K2_L2HINT=2)
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
for v in A B B2 A B B2; do
  hint=0; lib=$v
  if [ $v = B2 ]; then hint=2; lib=B; fi
  cp ftk_b200/libvariant$lib.so ftk_b200/libftkb200.so
  FTKB_K2_L2HINT=$hint timeout 200 python bench.py --config c2 --only-main --steps 60 --warmup 5 --e2e-steps 0 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02s2_c2_$v.json
  show gpurun_out/r02s2_c2_$v.json "c2 variant $v"
done
cp /tmp/lib_current.so ftk_b200/libftkb200.so
FTKB_K2_L2HINT=2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "2d or scalar" 2>&1 | tail -2
