#!/bin/bash
# round 2, GPU session 48 (1 GPU): 3D scalar build kernel reads the other layer's cells through a cp.async ring three planes ahead
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
for k in 1 2; do
  timeout 200 python bench.py --config c3 --only-main --steps 24 --warmup 4 --e2e-steps 0 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02x2_c3_$k.json
  show gpurun_out/r02x2_c3_$k.json "c3 run $k"
done
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2 | tee gpurun_out/r02x2_pytest_gpu.log
