#!/bin/bash
# round 2, GPU session 6: overlap probe; 3D scalar build kernel without calls in the plane loop
mkdir -p gpurun_out
./scripts/probe/overlap_probe 2>&1 | tee gpurun_out/r02f_overlap_probe.log
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r02f_pytest_gpu.log
run3() { name=$1; shift
  env FTKB_DEBUG_TIMING=1 "$@" timeout 600 python bench.py --config c3 --steps 31 --warmup 5 --no-cpu-baseline --e2e-steps 0 --only-main 2> gpurun_out/r02f_bench_$name.err | tee gpurun_out/r02f_bench_$name.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()}, round(d['roofline']['frac'],3))"
  grep ftkb gpurun_out/r02f_bench_$name.err | head -2
}
run3 c3 FTKB_X=1
run3 c3_rows32 FTKB_S3_ROWS=32
run3 c3_rows20 FTKB_S3_ROWS=20
timeout 600 python bench.py --config c5 --steps 40 --warmup 5 --no-cpu-baseline --e2e-steps 0 --only-main 2> gpurun_out/r02f_bench_c5.err | tee gpurun_out/r02f_bench_c5.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c5', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()}, round(d['roofline']['frac'],3), d['finalize_ms'], d['finalize_ms_device'], d['finalize_ms_host'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan3d_build -s 4 -c 1 -o gpurun_out/r02f_prof_c3 -f \
    python bench.py --config c3 --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 0 --only-main > gpurun_out/r02f_ncu_full.log 2>&1
ls -la gpurun_out | tail -3
