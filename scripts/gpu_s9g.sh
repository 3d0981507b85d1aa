#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
for rows in default 72 144; do
if [ $rows = default ]; then unset FTKB_C2_ROWS; else export FTKB_C2_ROWS=$rows; fi
timeout 300 python bench.py --steps 126 --no-cpu-baseline --e2e-steps 2 2> gpurun_out/bench_c2.err | tee gpurun_out/bench_c2_cells.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c2 rows=$rows', {k:d[k] for k in ('ms_per_step','kernel_ms_per_step')}, d['roofline']['frac'])"
done
