#!/bin/bash
# round 2, GPU session 50 (1 GPU): evidence refresh after the last kernel changes (3D cell ring): default bench line, C3 line, launch list and ncu summary of C3
T=r02zz
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/${T}_smoke.log
timeout 900 python bench.py 2> gpurun_out/${T}_bench_c2.err | tee gpurun_out/${T}_bench_c2.json | cut -c1-300
timeout 600 python bench.py --config c3 --steps 31 --only-main 2> gpurun_out/${T}_bench_c3.err | tee gpurun_out/${T}_bench_c3.json | cut -c1-300
K='regex:scan|test_kernel|gradient|resolution|fill_u64|synthetic|point_keys|gather_points|neighbors|uf_|Radix|Select|Unique'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 120 --csv --log-file gpurun_out/${T}_launches_c3.csv \
    python bench.py --config c3 --steps 12 --warmup 3 --no-cpu-baseline --e2e-steps 1 --only-main > gpurun_out/${T}_ncu_launch_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan3d_build -s 4 -c 1 -o gpurun_out/${T}_prof_c3 -f \
    python bench.py --config c3 --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 0 --only-main > gpurun_out/${T}_ncu_full_c3.log 2>&1
ls -la gpurun_out | grep ${T}_ | awk '{print $5, $9}'
