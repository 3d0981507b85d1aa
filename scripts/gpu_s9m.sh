#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
for mode in direct tile; do
FTKB_SCAN=$mode timeout 300 python bench.py --steps 126 --no-cpu-baseline --e2e-steps 2 2> gpurun_out/b.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c2 $mode', {k:d[k] for k in ('ms_per_step','kernel_ms_per_step','punctured_simplices','cells_refined_per_step')}, d['roofline']['frac'])"; tail -2 gpurun_out/b.err
done
for rows in 32 128; do
FTKB_C2_ROWS=$rows timeout 300 python bench.py --steps 126 --no-cpu-baseline --e2e-steps 2 2> gpurun_out/b.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c2 direct rows=$rows', {k:d[k] for k in ('ms_per_step','kernel_ms_per_step')}, d['roofline']['frac'])"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan2d_direct_kernel -s 4 -c 1 -o gpurun_out/prof_d2 -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_d2.log 2>&1
tail -1 gpurun_out/ncu_d2.log
