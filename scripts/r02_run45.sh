#!/bin/bash
# round 2, GPU session 45 (1 GPU): finalize phase laps on the dense field, with and without the e2e leg
mkdir -p gpurun_out
for e in 0 12; do
  FTKB_DEBUG_TIMING=1 timeout 300 python bench.py --config woven --only-main --steps 12 --warmup 3 --e2e-steps $e --no-cpu-baseline 2>gpurun_out/r02u2_woven_e$e.err | tail -1 > gpurun_out/r02u2_woven_e$e.json
  python - gpurun_out/r02u2_woven_e$e.json $e <<'P'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("e2e steps", sys.argv[2], "ms/step %.4f" % d["ms_per_step"], {k: d.get(k) for k in ("finalize_ms", "finalize_ms_library", "finalize_ms_device", "finalize_ms_host", "punctured_simplices")})
P
  grep "ftkb timing" gpurun_out/r02u2_woven_e$e.err | tail -3
done
