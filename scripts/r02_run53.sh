#!/bin/bash
# round 2, GPU session 53 (1 GPU): sorted records and keys inside the finalize slab (one cudaMalloc per finalize); full parity tests; default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2 | tee gpurun_out/r02ad2_pytest_gpu.log
timeout 900 python bench.py 2> gpurun_out/r02ad2_bench_c2.err | tee gpurun_out/r02ad2_bench_c2.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('c2', d['ms_per_step'], d['roofline']['frac'], d['finalize_ms'], d['finalize_ms_library'])
for k in ('scaling_c4','scaling_c3','dense_woven'):
    r=d[k]; print(k, r['value'], r['ms_per_step'], r.get('finalize_ms'), r.get('finalize_ms_library'))"
