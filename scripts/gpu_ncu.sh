#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:scan2d_bulk -s 4 -c 1 -o gpurun_out/prof_bulk -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_run.log 2>&1
tail -2 gpurun_out/ncu_full_run.log
