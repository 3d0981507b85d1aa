#!/bin/bash
# round 2, GPU session 33 (1 GPU): finalize phase laps on the dense configurations after the no-init host arrays
mkdir -p gpurun_out
for cfg in c5 woven; do
  FTKB_DEBUG_TIMING=1 timeout 300 python bench.py --config $cfg --only-main --steps 12 --warmup 3 --e2e-steps 0 --no-cpu-baseline 2>gpurun_out/r02j_${cfg}_dbg.err | tail -1 > gpurun_out/r02j_${cfg}_dbg.json
  python - gpurun_out/r02j_${cfg}_dbg.json $cfg <<'P'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], "ms/step %.4f" % d["ms_per_step"], {k: d.get(k) for k in ("finalize_ms", "finalize_ms_library", "finalize_ms_device", "finalize_ms_host", "punctured_simplices")})
P
  grep "ftkb timing" gpurun_out/r02j_${cfg}_dbg.err | tail -3
done
