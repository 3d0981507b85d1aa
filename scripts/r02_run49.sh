#!/bin/bash
# round 2, GPU session 49 (1 GPU): ncu capture of the 2D vector build kernel (C5) and of the current 3D scalar kernel
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vscan2d_build -s 4 -c 1 -o gpurun_out/r02y2_prof_c5 -f \
    python bench.py --config c5 --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 0 --only-main > gpurun_out/r02y2_ncu_c5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan3d_build -s 4 -c 1 -o gpurun_out/r02y2_prof_c3 -f \
    python bench.py --config c3 --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 0 --only-main > gpurun_out/r02y2_ncu_c3.log 2>&1
ls -la gpurun_out/r02y2_prof*.ncu-rep
