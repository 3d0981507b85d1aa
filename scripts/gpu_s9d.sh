#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
for rows in default 63 36 18; do
if [ $rows = default ]; then unset FTKB_C2_ROWS; else export FTKB_C2_ROWS=$rows; fi
timeout 300 python bench.py --steps 126 --no-cpu-baseline --e2e-steps 4 2> gpurun_out/bench_c2.err | tee gpurun_out/bench_c2_cells_$rows.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('rows=$rows', {k:d[k] for k in ('ms_per_step','kernel_ms_per_step','cells_refined_per_step','punctured_simplices')}, d['roofline']['frac'])"
tail -2 gpurun_out/bench_c2.err
done
unset FTKB_C2_ROWS
FTKB_SCAN2D=twolayer timeout 300 python bench.py --steps 126 --no-cpu-baseline --e2e-steps 4 2> gpurun_out/bench_c2.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('twolayer', {k:d[k] for k in ('ms_per_step','kernel_ms_per_step','cells_refined_per_step','punctured_simplices')}, d['roofline']['frac'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan2d_build -s 4 -c 1 -o gpurun_out/prof_c2build -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_run2.log 2>&1
tail -2 gpurun_out/ncu_full_run2.log
