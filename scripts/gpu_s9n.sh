#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
for i in 1 2; do
timeout 300 python bench.py --steps 252 --no-cpu-baseline --e2e-steps 2 2> gpurun_out/b.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c2', {k:d[k] for k in ('ms_per_step','kernel_ms_per_step','punctured_simplices','cells_refined_per_step')}, d['roofline']['frac'])"; tail -2 gpurun_out/b.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan2d_build -s 4 -c 1 -o gpurun_out/prof_c2build -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_run2.log 2>&1
