#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 126 --no-cpu-baseline --e2e-steps 2 2> gpurun_out/bench_c2.err | tee gpurun_out/bench_c2_cells.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c2 default', {k:d[k] for k in ('ms_per_step','kernel_ms_per_step')}, d['roofline']['frac'])"
for rows in default 64 32; do
if [ $rows = default ]; then unset FTKB_S3_ROWS; else export FTKB_S3_ROWS=$rows; fi
timeout 300 python bench.py --config c3 --steps 31 --no-cpu-baseline --e2e-steps 2 2> gpurun_out/bench_c3.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c3 rows=$rows', {k:d[k] for k in ('ms_per_step','kernel_ms_per_step')}, d['roofline']['frac'])"
done
unset FTKB_S3_ROWS
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan3d_build -s 4 -c 1 -o gpurun_out/prof_s3build -f \
    python bench.py --config c3 --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_run3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan2d_build -s 4 -c 1 -o gpurun_out/prof_c2build -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_run2.log 2>&1
