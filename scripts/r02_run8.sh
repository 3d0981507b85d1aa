#!/bin/bash
# round 2, GPU session 8: cell-pool fix (test kernels now really overlap the next scan), group tracker, wrap fixtures
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 | tee gpurun_out/r02h_pytest_gpu.log
run() { name=$1; cfg=$2; shift 2
  env FTKB_DEBUG_TIMING=1 "$@" timeout 600 python bench.py --config $cfg --steps 60 --warmup 10 --no-cpu-baseline --e2e-steps 0 --only-main 2> gpurun_out/r02h_bench_$name.err | tee gpurun_out/r02h_bench_$name.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()}, round(d['roofline']['frac'],3))"
  grep ftkb gpurun_out/r02h_bench_$name.err | head -3
}
run c2_k3 c2 FTKB_K2_CTAS=3
run c2_k2 c2 FTKB_K2_CTAS=2
run c2_k3_nooverlap c2 FTKB_TEST_OVERLAP=0
run c2_k3_carve c2 FTKB_CARVEOUT=1
run c3 c3 FTKB_X=1
run c5 c5 FTKB_X=1
run woven woven FTKB_X=1
timeout 900 python bench.py --steps 20 --warmup 5 2> gpurun_out/r02h_bench_default.err | tee gpurun_out/r02h_bench_default.json | cut -c1-200
tail -3 gpurun_out/r02h_bench_default.err
ls -la gpurun_out | tail -3
