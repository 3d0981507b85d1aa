#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "3d or 3D or abc or extremum" 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_3d.log
timeout 300 python bench.py --config c3 --steps 31 --no-cpu-baseline --e2e-steps 4 2> gpurun_out/bench_c3.err | tee gpurun_out/bench_c3_cells.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','kernel_ms_per_step','cells_refined_per_step')}, d['roofline']['frac'])"
tail -3 gpurun_out/bench_c3.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan3d_build -s 4 -c 1 -o gpurun_out/prof_s3build -f \
    python bench.py --config c3 --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_run3.log 2>&1
tail -2 gpurun_out/ncu_full_run3.log
