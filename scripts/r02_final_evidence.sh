#!/bin/bash
# round 2, final evidence session on one B200: parity tests, smoke, bench lines (C2 default with sub-records / CPU baseline / e2e, C3, C5,
# dense woven, reference arm), launch lists, ncu --set full captures of the two scalar build kernels and of the dense test kernel.
T=${TAG:-r02z}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/${T}_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/${T}_smoke.log
timeout 900 python bench.py 2> gpurun_out/${T}_bench_c2.err | tee gpurun_out/${T}_bench_c2.json | cut -c1-400
timeout 600 python bench.py --config c3 --steps 31 --only-main 2> gpurun_out/${T}_bench_c3.err | tee gpurun_out/${T}_bench_c3.json | cut -c1-300
timeout 600 python bench.py --config c5 --steps 60 --e2e-steps 4 --only-main 2> gpurun_out/${T}_bench_c5.err | tee gpurun_out/${T}_bench_c5.json | cut -c1-300
timeout 600 python bench.py --config woven --steps 12 --only-main 2> gpurun_out/${T}_bench_woven.err | tee gpurun_out/${T}_bench_woven.json | cut -c1-300
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2> gpurun_out/${T}_bench_ref.err | tee gpurun_out/${T}_bench_reference_arm.json | cut -c1-300
K='regex:scan|test_kernel|gradient|resolution|fill_u64|synthetic|point_keys|gather_points|neighbors|uf_|Radix|Select|Unique'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 120 --csv --log-file gpurun_out/${T}_launches_c2.csv \
    python bench.py --steps 12 --warmup 3 --no-cpu-baseline --e2e-steps 1 --only-main > gpurun_out/${T}_ncu_launch_c2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 120 --csv --log-file gpurun_out/${T}_launches_c3.csv \
    python bench.py --config c3 --steps 12 --warmup 3 --no-cpu-baseline --e2e-steps 1 --only-main > gpurun_out/${T}_ncu_launch_c3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 120 --csv --log-file gpurun_out/${T}_launches_woven.csv \
    python bench.py --config woven --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 0 --only-main > gpurun_out/${T}_ncu_launch_woven.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan2d_keys_build -s 4 -c 1 -o gpurun_out/${T}_prof_c2keys -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 0 --only-main > gpurun_out/${T}_ncu_full_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan3d_build -s 4 -c 1 -o gpurun_out/${T}_prof_c3 -f \
    python bench.py --config c3 --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 0 --only-main > gpurun_out/${T}_ncu_full_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:test_kernel -s 4 -c 1 -o gpurun_out/${T}_prof_woven_test -f \
    python bench.py --config woven --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 0 --only-main > gpurun_out/${T}_ncu_full_woven.log 2>&1
ls -la gpurun_out | grep ${T}_ | awk '{print $5, $9}'
