#!/bin/bash
# round 2, GPU session 4: persistent key kernel (no calls in the hot loop, cells prefetched), gap diagnostic with medians
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r02d_pytest_gpu.log
run() { # name, env...
  name=$1; shift
  env FTKB_DEBUG_TIMING=1 "$@" timeout 600 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --e2e-steps 4 --only-main 2> gpurun_out/r02d_bench_$name.err | tee gpurun_out/r02d_bench_$name.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'])"
  grep ftkb gpurun_out/r02d_bench_$name.err | head -3
}
run c2_keys FTKB_X=1
run c2_keys_carve FTKB_CARVEOUT=1
run c2_keys_nooverlap FTKB_TEST_OVERLAP=0
run c2_f32 FTKB_SCAN2D=f32
timeout 900 python bench.py --config woven --steps 12 --warmup 3 --no-cpu-baseline --e2e-steps 0 --only-main 2> gpurun_out/r02d_bench_woven.err | tee gpurun_out/r02d_bench_woven.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'], d['trajectories'], d['punctured_simplices'], d['finalize_ms'])"
FTKB_SCAN2D=f32 timeout 900 python bench.py --config woven --steps 12 --warmup 3 --no-cpu-baseline --e2e-steps 0 --only-main 2> gpurun_out/r02d_bench_woven_f32.err | tee gpurun_out/r02d_bench_woven_f32.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'], d['trajectories'], d['punctured_simplices'], d['finalize_ms'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan2d_keys_build -s 4 -c 1 -o gpurun_out/r02d_prof_c2keys -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 --only-main > gpurun_out/r02d_ncu_full.log 2>&1
ls -la gpurun_out | tail -3
