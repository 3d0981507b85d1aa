#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
K='regex:scan|test_kernel|gradient|resolution|fill_u64|synthetic|point_keys|gather_points|neighbors|uf_|Radix|Select|Unique'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 120 --csv --log-file gpurun_out/launches_c2.csv \
    python bench.py --steps 12 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch_run.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 120 --csv --log-file gpurun_out/launches_c3.csv \
    python bench.py --config c3 --steps 12 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch_run3.log 2>&1
tail -2 gpurun_out/ncu_launch_run.log
