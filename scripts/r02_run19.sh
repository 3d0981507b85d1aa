#!/bin/bash
# round 2, GPU session 19 (1 GPU): chunk-length sweeps of the 2D key kernel and the 3D build kernel (waves of CTAs),
# finalize phase laps on the dense config, CLI input timing with its split, float32 test
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_cli.py -m gpu -q -x -k "float32 or raw_input" 2>&1 | tail -3
show() { python - "$1" "$2" <<'P'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], "ms/step %.4f scan %.4f frac %.3f" % (d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"]))
P
}
for r in 27 36 45 54 63 72 90 126; do
  FTKB_C2_ROWS=$r timeout 200 python bench.py --config c2 --only-main --steps 60 --warmup 5 --e2e-steps 0 2>/dev/null | tail -1 > gpurun_out/r02s_c2_rows$r.json
  show gpurun_out/r02s_c2_rows$r.json "c2 rows $r"
done
for r in 45 63; do
  FTKB_K2_CTAS=4 FTKB_C2_ROWS=$r timeout 200 python bench.py --config c2 --only-main --steps 60 --warmup 5 --e2e-steps 0 2>/dev/null | tail -1 > gpurun_out/r02s_c2_ctas4_rows$r.json
  show gpurun_out/r02s_c2_ctas4_rows$r.json "c2 4 CTAs/SM rows $r"
done
for r in 32 40 43 52; do
  FTKB_S3_ROWS=$r timeout 200 python bench.py --config c3 --only-main --steps 24 --warmup 4 --e2e-steps 0 2>/dev/null | tail -1 > gpurun_out/r02s_c3_rows$r.json
  show gpurun_out/r02s_c3_rows$r.json "c3 rows $r"
done
FTKB_DEBUG_TIMING=1 timeout 300 python bench.py --config woven --only-main --steps 12 --warmup 3 --e2e-steps 0 2> gpurun_out/r02s_woven_timing.err | tail -1 > gpurun_out/r02s_bench_woven.json
grep "ftkb timing" gpurun_out/r02s_woven_timing.err | tail -6
TAG=r02s timeout 400 bash scripts/cli_input_timing.sh 2>&1 | tail -14
