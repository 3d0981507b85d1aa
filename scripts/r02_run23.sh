#!/bin/bash
# round 2, GPU session 23 (1 GPU): persistent 2D key kernel with a device work counter (FTKB_K2_PERSIST=1) against the grid-per-chunk one
mkdir -p gpurun_out
FTKB_K2_PERSIST=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_wrap.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r02w_pytest_persist.log
show() { python - "$1" "$2" <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms/step %.4f scan %.4f frac %.3f value %.3e" % (d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["value"]), d.get("kernel_ms_per_step"), d.get("trajectories"), d.get("punctured_simplices"))
except Exception as e:
    print(sys.argv[2], "failed", e)
P
}
timeout 200 python bench.py --config c2 --only-main --steps 60 --warmup 5 --e2e-steps 0 2>/dev/null | tail -1 > gpurun_out/r02w_c2_grid.json
show gpurun_out/r02w_c2_grid.json "c2 grid-per-chunk rows 45"
for r in 18 27 36 45 63; do
  FTKB_K2_PERSIST=1 FTKB_C2_ROWS=$r timeout 200 python bench.py --config c2 --only-main --steps 60 --warmup 5 --e2e-steps 0 2>gpurun_out/r02w_c2_persist_rows$r.err | tail -1 > gpurun_out/r02w_c2_persist_rows$r.json
  show gpurun_out/r02w_c2_persist_rows$r.json "c2 persistent rows $r"
done
FTKB_K2_PERSIST=1 timeout 300 python bench.py --config woven --only-main --steps 12 --warmup 3 --e2e-steps 0 2>/dev/null | tail -1 > gpurun_out/r02w_woven_persist.json
show gpurun_out/r02w_woven_persist.json "woven persistent"
timeout 300 python bench.py --config woven --only-main --steps 12 --warmup 3 --e2e-steps 0 2>/dev/null | tail -1 > gpurun_out/r02w_woven_grid.json
show gpurun_out/r02w_woven_grid.json "woven grid-per-chunk"
