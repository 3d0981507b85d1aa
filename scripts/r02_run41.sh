#!/bin/bash
# round 2, GPU session 41 (1 GPU): 2D key kernel stages the other layer's cell with cp.async
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
  timeout 200 python bench.py --config c2 --only-main --steps 60 --warmup 5 --e2e-steps 0 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02r2_c2_$k.json
  show gpurun_out/r02r2_c2_$k.json "c2 run $k"
done
for r in 36 54; do
  FTKB_C2_ROWS=$r timeout 200 python bench.py --config c2 --only-main --steps 60 --warmup 5 --e2e-steps 0 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02r2_c2_rows$r.json
  show gpurun_out/r02r2_c2_rows$r.json "c2 rows $r"
done
timeout 300 python bench.py --config woven --only-main --steps 12 --warmup 3 --e2e-steps 0 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02r2_woven.json
show gpurun_out/r02r2_woven.json "woven"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_wrap.py tests/test_pyftk.py -m gpu -q -x 2>&1 | tail -2
