#!/bin/bash
# round 2, GPU session 9: ring depth / chunk size of the key kernel, CLI file-series e2e, pooled finalize, full default bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/r02i_pytest_gpu.log
run() { name=$1; cfg=$2; shift 2
  env "$@" timeout 600 python bench.py --config $cfg --steps 60 --warmup 10 --no-cpu-baseline --e2e-steps 0 --only-main 2> gpurun_out/r02i_bench_$name.err | tee gpurun_out/r02i_bench_$name.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()}, round(d['roofline']['frac'],3), d['finalize_ms'], d['finalize_ms_device'], d['finalize_ms_host'])"
}
run c2_nst4 c2 FTKB_K2_NST=4
run c2_nst5 c2 FTKB_K2_NST=5
run c2_nst6 c2 FTKB_K2_NST=6
run c2_nst6_rows54 c2 FTKB_K2_NST=6 FTKB_C2_ROWS=54
run c2_nst6_rows81 c2 FTKB_K2_NST=6 FTKB_C2_ROWS=81
run c2_nst4_rows54 c2 FTKB_K2_NST=4 FTKB_C2_ROWS=54
run c2_nst4_rows81 c2 FTKB_K2_NST=4 FTKB_C2_ROWS=81
run c5 c5 FTKB_X=1
run woven woven FTKB_X=1
bash scripts/cli_input_timing.sh gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 2> gpurun_out/r02i_bench_default.err | tee gpurun_out/r02i_bench_default.json | cut -c1-200
tail -3 gpurun_out/r02i_bench_default.err
ls -la gpurun_out | tail -3
