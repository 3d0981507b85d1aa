#!/bin/bash
# round 2, GPU session 1: deferred (sync-free) steps -- parity tests, then bench A/B with FTKB_DEFER=0
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r02a_pytest_gpu.log
for d in 1 0; do
  FTKB_DEFER=$d timeout 600 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --e2e-steps 4 2> gpurun_out/r02a_bench_c2_defer$d.err | tee gpurun_out/r02a_bench_c2_defer$d.json | cut -c1-400
  FTKB_DEFER=$d timeout 600 python bench.py --config c3 --steps 31 --no-cpu-baseline --e2e-steps 2 2> gpurun_out/r02a_bench_c3_defer$d.err | tee gpurun_out/r02a_bench_c3_defer$d.json | cut -c1-400
done
FTKB_DEFER=1 timeout 600 python bench.py --config c5 --steps 60 --no-cpu-baseline --e2e-steps 2 2> gpurun_out/r02a_bench_c5.err | tee gpurun_out/r02a_bench_c5.json | cut -c1-400
K='regex:scan|test_kernel'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 60 --csv --log-file gpurun_out/r02a_launches_c2.csv \
    python bench.py --steps 12 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r02a_ncu_launch_run.log 2>&1
tail -n 30 gpurun_out/r02a_launches_c2.csv | cut -c1-200
