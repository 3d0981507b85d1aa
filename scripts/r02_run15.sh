#!/bin/bash
# round 2, GPU session 15: z-slab groups (spatial decomposition), test kernel back to one phase
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_group.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r02o_pytest_group.log
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/r02o_pytest_gpu.log
run() { name=$1; cfg=$2; shift 2
  env "$@" timeout 600 python bench.py --config $cfg --steps 40 --warmup 6 --no-cpu-baseline --e2e-steps 0 --only-main 2> gpurun_out/r02o_bench_$name.err | tee gpurun_out/r02o_bench_$name.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()}, round(d['roofline']['frac'],3), d['finalize_ms'], d['finalize_ms_library'])"
}
run c2 c2 FTKB_X=1
run woven woven FTKB_X=1
run woven_nooverlap woven FTKB_TEST_OVERLAP=0
ls -la gpurun_out | tail -3
