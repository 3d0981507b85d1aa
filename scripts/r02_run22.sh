#!/bin/bash
# round 2, GPU session 22 (1 GPU): test kernel in three stages with shared-memory queues (cheap exclusion -> exact predicate -> record)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r02v_pytest_gpu.log
show() { python - "$1" "$2" <<'P'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], "ms/step %.4f scan %.4f frac %.3f value %.3e" % (d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["value"]), d.get("kernel_ms_per_step"))
P
}
timeout 300 python bench.py --config woven --only-main --steps 12 --warmup 3 --e2e-steps 0 2>/dev/null | tail -1 > gpurun_out/r02v_woven.json
show gpurun_out/r02v_woven.json "woven"
FTKB_TEST_OVERLAP=0 timeout 300 python bench.py --config woven --only-main --steps 12 --warmup 3 --e2e-steps 0 2>/dev/null | tail -1 > gpurun_out/r02v_woven_nooverlap.json
show gpurun_out/r02v_woven_nooverlap.json "woven, test kernel on the sweep's stream"
timeout 200 python bench.py --config c2 --only-main --steps 60 --warmup 5 --e2e-steps 0 2>/dev/null | tail -1 > gpurun_out/r02v_c2.json
show gpurun_out/r02v_c2.json "c2"
timeout 200 python bench.py --config c3 --only-main --steps 24 --warmup 4 --e2e-steps 0 2>/dev/null | tail -1 > gpurun_out/r02v_c3.json
show gpurun_out/r02v_c3.json "c3"
timeout 300 python bench.py --config c5 --only-main --steps 30 --warmup 4 --e2e-steps 0 2>/dev/null | tail -1 > gpurun_out/r02v_c5.json
show gpurun_out/r02v_c5.json "c5"
