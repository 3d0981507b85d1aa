#!/bin/bash
# round 2, GPU session 14: two-phase test kernel (pretest, compaction, full test), packed cold path; CPU baselines of the vector configurations
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/r02n_pytest_gpu.log
run() { name=$1; cfg=$2; shift 2
  env "$@" timeout 600 python bench.py --config $cfg --steps 40 --warmup 6 --no-cpu-baseline --e2e-steps 0 --only-main 2> gpurun_out/r02n_bench_$name.err | tee gpurun_out/r02n_bench_$name.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()}, round(d['roofline']['frac'],3), d['finalize_ms'], d['finalize_ms_library'])"
}
run c2 c2 FTKB_X=1
run c3 c3 FTKB_X=1
run woven woven FTKB_X=1
run woven_nooverlap woven FTKB_TEST_OVERLAP=0
timeout 900 python bench.py --config c5 --steps 40 --e2e-steps 4 --only-main 2> gpurun_out/r02n_bench_c5_full.err | tee gpurun_out/r02n_bench_c5_full.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c5', d['ms_per_step'], d['value'], d['cpu_baseline'], d['e2e'])"
timeout 900 python bench.py --config c4 --steps 12 --e2e-steps 0 --only-main 2> gpurun_out/r02n_bench_c4_full.err | tee gpurun_out/r02n_bench_c4_full.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c4', d['ms_per_step'], d['value'], d['cpu_baseline'])"
timeout 900 python bench.py --config c3 --steps 31 --e2e-steps 4 --only-main 2> gpurun_out/r02n_bench_c3_full.err | tee gpurun_out/r02n_bench_c3_full.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c3', d['ms_per_step'], d['value'], d['cpu_baseline'], d['e2e'])"
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/r02n_smoke.log
ls -la gpurun_out | tail -3
