#!/bin/bash
# fused 3D scan: parity first, then C3 bench + ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "random or fused or golden or 3d" 2>&1 | tail -15 | tee gpurun_out/pytest_f3.log
timeout 300 python bench.py --config c3 --steps 31 --no-cpu-baseline --e2e-steps 4 2> gpurun_out/bench_c3.err | tee gpurun_out/bench_f3_c3.json | cut -c1-300
tail -3 gpurun_out/bench_c3.err
python -c "
import json;d=json.load(open('gpurun_out/bench_f3_c3.json'));print(d['kernel_ms_per_step'], d['roofline']['frac'], d['cells_refined_per_step'], d['ms_per_step'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan3d_fused -s 4 -c 1 -o gpurun_out/prof_f3 -f \
    python bench.py --config c3 --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_f3_run.log 2>&1
tail -2 gpurun_out/ncu_f3_run.log
