#!/bin/bash
# round 2, GPU session 35 (1 GPU): 2D key kernel with an L2 evict-first policy on the layer's bulk copies
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
for h in 0 1 0 1; do
  FTKB_K2_L2HINT=$h timeout 200 python bench.py --config c2 --only-main --steps 60 --warmup 5 --e2e-steps 0 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02l_c2_hint$h.json
  show gpurun_out/r02l_c2_hint$h.json "c2 l2 hint $h"
done
FTKB_K2_L2HINT=1 timeout 300 python bench.py --config woven --only-main --steps 12 --warmup 3 --e2e-steps 0 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02l_woven_hint1.json
show gpurun_out/r02l_woven_hint1.json "woven l2 hint 1"
FTKB_K2_L2HINT=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "2d or scalar" 2>&1 | tail -2
