#!/bin/bash
# round 2, GPU session 13: compute-sanitizer racecheck / synccheck / memcheck on the ring kernels; final ncu captures
mkdir -p gpurun_out
for tool in racecheck synccheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_rings.py > gpurun_out/r02m_sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; tail -4 gpurun_out/r02m_sanitizer_$tool.log
done
K='regex:scan|test_kernel'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 80 --csv --log-file gpurun_out/r02m_launches_c2.csv \
    python bench.py --steps 12 --warmup 3 --no-cpu-baseline --e2e-steps 1 --only-main > gpurun_out/r02m_ncu_launch_run.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 80 --csv --log-file gpurun_out/r02m_launches_c3.csv \
    python bench.py --config c3 --steps 12 --warmup 3 --no-cpu-baseline --e2e-steps 1 --only-main > gpurun_out/r02m_ncu_launch_run3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan2d_keys_build -s 4 -c 1 -o gpurun_out/r02m_prof_c2keys -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 0 --only-main > gpurun_out/r02m_ncu_full_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan3d_build -s 4 -c 1 -o gpurun_out/r02m_prof_c3 -f \
    python bench.py --config c3 --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 0 --only-main > gpurun_out/r02m_ncu_full_c3.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:test_kernel -s 4 -c 1 -o gpurun_out/r02m_prof_woven_test -f \
    python bench.py --config woven --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 0 --only-main > gpurun_out/r02m_ncu_full_woven.log 2>&1
ls -la gpurun_out | tail -4
