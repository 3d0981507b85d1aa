#!/bin/bash
# round 2, GPU session 24 (1 GPU): point buffer grown ahead of the deferred steps (no overflow replays on feature-dense fields)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r02x_pytest_gpu.log
show() { python - "$1" "$2" <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms/step %.4f scan %.4f frac %.3f value %.3e" % (d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["value"]), d.get("kernel_ms_per_step"), "repeated", d["roofline"].get("sweeps_repeated"), d.get("trajectories"), d.get("punctured_simplices"))
except Exception as e:
    print(sys.argv[2], "failed", e)
P
}
for k in 1 2 3; do
  timeout 300 python bench.py --config woven --only-main --steps 12 --warmup 3 --e2e-steps 0 2>/dev/null | tail -1 > gpurun_out/r02x_woven_$k.json
  show gpurun_out/r02x_woven_$k.json "woven run $k"
done
FTKB_DEBUG_TIMING=1 timeout 300 python bench.py --config woven --only-main --steps 12 --warmup 3 --e2e-steps 0 2>gpurun_out/r02x_woven_dbg.err | tail -1 > gpurun_out/r02x_woven_dbg.json
show gpurun_out/r02x_woven_dbg.json "woven, debug timing"
grep "ftkb" gpurun_out/r02x_woven_dbg.err | tail -8
timeout 300 python bench.py --config woven --only-main --steps 30 --warmup 3 --e2e-steps 0 2>/dev/null | tail -1 > gpurun_out/r02x_woven_30.json
show gpurun_out/r02x_woven_30.json "woven 30 steps"
timeout 200 python bench.py --config c2 --only-main --steps 60 --warmup 5 --e2e-steps 0 2>/dev/null | tail -1 > gpurun_out/r02x_c2.json
show gpurun_out/r02x_c2.json "c2"
