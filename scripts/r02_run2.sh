#!/bin/bash
# round 2, GPU session 2: key-based 2D build kernel + test kernels overlapped on a second stream -- parity tests, bench A/B
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r02b_pytest_gpu.log
run() { # name, env..., -- args
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --e2e-steps 4 2> gpurun_out/r02b_bench_$name.err | tee gpurun_out/r02b_bench_$name.json | cut -c1-330
}
run c2_keys FTKB_X=1
run c2_f32 FTKB_SCAN2D=f32
run c2_keys_nooverlap FTKB_TEST_OVERLAP=0
run c2_keys_sync FTKB_DEFER=0
timeout 600 python bench.py --config c3 --steps 31 --no-cpu-baseline --e2e-steps 2 2> gpurun_out/r02b_bench_c3.err | tee gpurun_out/r02b_bench_c3.json | cut -c1-330
timeout 600 python bench.py --config c5 --steps 60 --no-cpu-baseline --e2e-steps 2 2> gpurun_out/r02b_bench_c5.err | tee gpurun_out/r02b_bench_c5.json | cut -c1-330
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan2d_keys_build -s 4 -c 1 -o gpurun_out/r02b_prof_c2keys -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r02b_ncu_full.log 2>&1
ls -la gpurun_out | tail -5
