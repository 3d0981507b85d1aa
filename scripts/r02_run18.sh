#!/bin/bash
# round 2, GPU session 18 (1 GPU): full GPU suite on the library with the threaded ordering walk, vector cells built by the
# resolution pass and float32 pushes; finalize timing on the dense configs; float32 CLI timing; C3 chunk-length A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/r02r_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/r02r_smoke.log
timeout 300 python bench.py --config woven --only-main --steps 12 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r02r_bench_woven.json; cut -c1-600 gpurun_out/r02r_bench_woven.json
timeout 300 python bench.py --config c5 --only-main --steps 12 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r02r_bench_c5.json; cut -c1-600 gpurun_out/r02r_bench_c5.json
for r in 40 63; do
  FTKB_S3_ROWS=$r timeout 300 python bench.py --config c3 --only-main --steps 12 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r02r_bench_c3_rows$r.json
  python - <<P
import json
d=json.loads(open("gpurun_out/r02r_bench_c3_rows$r.json").read())
print("c3 rows $r", d["ms_per_step"], d["roofline"]["frac"], d.get("phases"))
P
done
TAG=r02r timeout 400 bash scripts/cli_input_timing.sh 2>&1 | tail -12
