#!/bin/bash
# one GPU session: parity tests, bench, ncu launch list, ncu full capture of the scan kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python bench.py 2> gpurun_out/bench.err | tee gpurun_out/bench.json
tail -3 gpurun_out/bench.err
python bench.py --config c3 --steps 31 --no-cpu-baseline 2> gpurun_out/bench_c3.err | tee gpurun_out/bench_c3.json
tail -3 gpurun_out/bench_c3.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan2d_fused -s 4 -c 2 -o gpurun_out/prof_scan2d \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_run.log 2>&1
ls -la gpurun_out
