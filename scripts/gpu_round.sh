#!/bin/bash
# one evidence session on a B200 box: parity tests, smoke, bench lines (C2 default with the CPU baseline, C3, C5, C4, reference arm),
# launch lists of the library's kernels, ncu --set full captures of the four build kernels.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 600 python bench.py 2> gpurun_out/bench.err | tee gpurun_out/bench_c2.json | cut -c1-300
timeout 600 python bench.py --config c3 --steps 31 --no-cpu-baseline 2> gpurun_out/bench_c3.err | tee gpurun_out/bench_c3.json | cut -c1-300
timeout 600 python bench.py --config c5 --steps 60 --e2e-steps 4 2> gpurun_out/bench_c5.err | tee gpurun_out/bench_c5.json | cut -c1-300
timeout 900 python bench.py --config c4 --steps 12 --e2e-steps 0 2> gpurun_out/bench_c4.err | tee gpurun_out/bench_c4.json | cut -c1-300
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2> gpurun_out/bench_ref.err | tee gpurun_out/bench_ref.json | cut -c1-300
K='regex:scan|test_kernel|gradient|resolution|fill_u64|synthetic|point_keys|gather_points|neighbors|uf_|Radix|Select|Unique'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 120 --csv --log-file gpurun_out/launches_c2.csv \
    python bench.py --steps 12 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch_run.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 120 --csv --log-file gpurun_out/launches_c3.csv \
    python bench.py --config c3 --steps 12 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch_run3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan2d_build -s 4 -c 1 -o gpurun_out/prof_c2build -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_run2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan3d_build -s 4 -c 1 -o gpurun_out/prof_s3build -f \
    python bench.py --config c3 --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_run3.log 2>&1
ls -la gpurun_out | tail -20
