#!/bin/bash
# round 2, GPU session 7: key kernel with filtered exact minima / full-group fast path; timeline diagnostic; default bench with sub-records
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r02g_pytest_gpu.log
run() { name=$1; cfg=$2; shift 2
  env FTKB_DEBUG_TIMING=1 "$@" timeout 600 python bench.py --config $cfg --steps 60 --warmup 10 --no-cpu-baseline --e2e-steps 0 --only-main 2> gpurun_out/r02g_bench_$name.err | tee gpurun_out/r02g_bench_$name.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()}, round(d['roofline']['frac'],3))"
  grep ftkb gpurun_out/r02g_bench_$name.err | head -3
}
run c2_k3 c2 FTKB_K2_CTAS=3
run c2_k2 c2 FTKB_K2_CTAS=2
run c2_k3_nooverlap c2 FTKB_TEST_OVERLAP=0
run c2_k3_sync c2 FTKB_DEFER=0
run c3 c3 FTKB_X=1
timeout 900 python bench.py --steps 20 --warmup 5 2> gpurun_out/r02g_bench_default.err | tee gpurun_out/r02g_bench_default.json | cut -c1-200
tail -3 gpurun_out/r02g_bench_default.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2> gpurun_out/r02g_bench_ref.err | tee gpurun_out/r02g_bench_ref.json | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan2d_keys_build -s 4 -c 1 -o gpurun_out/r02g_prof_c2keys -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 0 --only-main > gpurun_out/r02g_ncu_full.log 2>&1
ls -la gpurun_out | tail -3
