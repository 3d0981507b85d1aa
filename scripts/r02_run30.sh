#!/bin/bash
# round 2, GPU session 30 (1 GPU): software-pipelined gather in the test kernel; then the sanitizers of session 29
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r02ab_pytest_gpu.log
show() { python - "$1" "$2" <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms/step %.4f scan %.4f frac %.3f value %.3e" % (d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["value"]), d.get("kernel_ms_per_step"), "repeated", d["roofline"].get("sweeps_repeated"), d.get("trajectories"), d.get("punctured_simplices"))
except Exception as e:
    print(sys.argv[2], "failed", e)
P
}
for k in 1 2; do
  timeout 300 python bench.py --config woven --only-main --steps 12 --warmup 3 --e2e-steps 0 2>/dev/null | tail -1 > gpurun_out/r02ab_woven_$k.json
  show gpurun_out/r02ab_woven_$k.json "woven run $k"
done
timeout 200 python bench.py --config c2 --only-main --steps 60 --warmup 5 --e2e-steps 0 2>/dev/null | tail -1 > gpurun_out/r02ab_c2.json
show gpurun_out/r02ab_c2.json "c2"
timeout 200 python bench.py --config c3 --only-main --steps 24 --warmup 4 --e2e-steps 0 2>/dev/null | tail -1 > gpurun_out/r02ab_c3.json
show gpurun_out/r02ab_c3.json "c3"
timeout 300 python bench.py --config c5 --only-main --steps 30 --warmup 4 --e2e-steps 0 2>/dev/null | tail -1 > gpurun_out/r02ab_c5.json
show gpurun_out/r02ab_c5.json "c5"
bash scripts/r02_run29.sh
