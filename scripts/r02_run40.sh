#!/bin/bash
# round 2, GPU session 40 (1 GPU): ncu captures of the current 3D and 2D scalar build kernels (next hot spots)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan3d_build -s 4 -c 1 -o gpurun_out/r02q2_prof_c3 -f \
    python bench.py --config c3 --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 0 --only-main > gpurun_out/r02q2_ncu_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan2d_keys_build -s 4 -c 1 -o gpurun_out/r02q2_prof_c2keys -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 0 --only-main > gpurun_out/r02q2_ncu_c2.log 2>&1
ls -la gpurun_out/r02q2_prof*.ncu-rep
