#!/bin/bash
# session 9 evidence run: parity tests, bench lines (C2 default, C3, reference arm), ncu launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py 2> gpurun_out/bench.err | tee gpurun_out/bench_c2.json
tail -3 gpurun_out/bench.err
timeout 600 python bench.py --config c3 --steps 31 --no-cpu-baseline 2> gpurun_out/bench_c3.err | tee gpurun_out/bench_c3.json
tail -3 gpurun_out/bench_c3.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2> gpurun_out/bench_ref.err | tee gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c2.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch_run.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c3.csv \
    python bench.py --config c3 --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch_run3.log 2>&1
ls -la gpurun_out | tail -12
