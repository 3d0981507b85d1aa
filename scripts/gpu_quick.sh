#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
for mode in tile warp; do
FTKB_SCAN=$mode timeout 300 python bench.py --steps 60 --no-cpu-baseline --e2e-steps 4 2> gpurun_out/bench.err | tee gpurun_out/bench_quick_$mode.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$mode', {k:d[k] for k in ('value','ms_per_step','kernel_ms_per_step','cells_refined_per_step')}, d['roofline']['frac'])"
tail -3 gpurun_out/bench.err
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan2d_tile -s 4 -c 1 -o gpurun_out/prof_tile -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_run.log 2>&1
tail -2 gpurun_out/ncu_full_run.log
