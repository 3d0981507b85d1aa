#!/bin/bash
# round 2, GPU session 3: key kernel with branch-free exact min |v|, carve-outs, gap diagnostic; new bench.py with sub-records
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r02c_pytest_gpu.log
run() { # name, env...
  name=$1; shift
  env FTKB_DEBUG_TIMING=1 "$@" timeout 600 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --e2e-steps 4 --only-main 2> gpurun_out/r02c_bench_$name.err | tee gpurun_out/r02c_bench_$name.json | cut -c1-250
  grep ftkb gpurun_out/r02c_bench_$name.err | head -3
}
run c2_keys FTKB_X=1
run c2_keys_nocarve FTKB_CARVEOUT=0
run c2_keys_nooverlap FTKB_TEST_OVERLAP=0
run c2_f32 FTKB_SCAN2D=f32
run c2_f32_nooverlap FTKB_SCAN2D=f32 FTKB_TEST_OVERLAP=0
timeout 900 python bench.py --steps 20 --warmup 5 2> gpurun_out/r02c_bench_default.err | tee gpurun_out/r02c_bench_default.json | cut -c1-250
tail -5 gpurun_out/r02c_bench_default.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan2d_keys_build -s 4 -c 1 -o gpurun_out/r02c_prof_c2keys -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 --only-main > gpurun_out/r02c_ncu_full.log 2>&1
ls -la gpurun_out | tail -4
