#!/bin/bash
# round 2, GPU session 5: key kernel at 2 / 3 / 4 CTAs per SM, chunk sizes, host-side timing of the deferred steps, woven
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r02e_pytest_gpu.log
run() { # name, env...
  name=$1; shift
  env FTKB_DEBUG_TIMING=1 "$@" timeout 600 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --e2e-steps 0 --only-main 2> gpurun_out/r02e_bench_$name.err | tee gpurun_out/r02e_bench_$name.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()}, round(d['roofline']['frac'],3))"
  grep ftkb gpurun_out/r02e_bench_$name.err | head -2
}
run c2_k3 FTKB_K2_CTAS=3
run c2_k2 FTKB_K2_CTAS=2
run c2_k4 FTKB_K2_CTAS=4
run c2_k3_rows36 FTKB_K2_CTAS=3 FTKB_C2_ROWS=36
run c2_k3_rows126 FTKB_K2_CTAS=3 FTKB_C2_ROWS=126
run c2_k3_rows252 FTKB_K2_CTAS=3 FTKB_C2_ROWS=252
run c2_k3_nooverlap FTKB_K2_CTAS=3 FTKB_TEST_OVERLAP=0
for m in keys f32; do
FTKB_SCAN2D=$m timeout 900 python bench.py --config woven --steps 12 --warmup 3 --no-cpu-baseline --e2e-steps 0 --only-main 2> gpurun_out/r02e_bench_woven_$m.err | tee gpurun_out/r02e_bench_woven_$m.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('woven $m', d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'], d['trajectories'], d['punctured_simplices'], d['finalize_ms'], d['finalize_ms_device'], d['finalize_ms_host'])"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan2d_keys_build -s 4 -c 1 -o gpurun_out/r02e_prof_c2keys -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 0 --only-main > gpurun_out/r02e_ncu_full.log 2>&1
ls -la gpurun_out | tail -3
