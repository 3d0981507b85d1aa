#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 120 python scripts/dbg_wl.py 2>&1 | tail -8
timeout 300 python bench.py --steps 60 --no-cpu-baseline --e2e-steps 4 2> gpurun_out/bench.err | tee gpurun_out/bench_quick.json
tail -3 gpurun_out/bench.err
FTKB_SCAN=ldg timeout 300 python bench.py --steps 60 --no-cpu-baseline --e2e-steps 4 2> gpurun_out/bench.err | tee gpurun_out/bench_quick_ldg.json
