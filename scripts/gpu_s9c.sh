#!/bin/bash
mkdir -p gpurun_out
for rows in 256 128 64 32 16; do
FTKB_S3_ROWS=$rows timeout 300 python bench.py --config c3 --steps 31 --no-cpu-baseline --e2e-steps 2 2> gpurun_out/bench_c3.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('rows=$rows', {k:d[k] for k in ('ms_per_step','kernel_ms_per_step')}, d['roofline']['frac'])"
done
