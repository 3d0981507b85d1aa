#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 300 python bench.py --config c4s --steps 20 --e2e-steps 2 2> gpurun_out/b.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c4s', {k:d[k] for k in ('ms_per_step','kernel_ms_per_step','punctured_simplices','cells_refined_per_step')}, d['roofline']['frac'])"; tail -2 gpurun_out/b.err
timeout 600 python bench.py --config c5 --steps 60 --e2e-steps 4 2> gpurun_out/b.err | tee gpurun_out/bench_c5.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c5', {k:d[k] for k in ('ms_per_step','kernel_ms_per_step','punctured_simplices','cells_refined_per_step')}, d['roofline']['frac'])"; tail -2 gpurun_out/b.err
timeout 900 python bench.py --config c4 --steps 12 --e2e-steps 0 2> gpurun_out/b.err | tee gpurun_out/bench_c4.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c4', {k:d[k] for k in ('ms_per_step','kernel_ms_per_step','punctured_simplices','cells_refined_per_step')}, d['roofline']['frac'])"; tail -2 gpurun_out/b.err
FTKB_VSCAN=twolayer timeout 600 python bench.py --config c5 --steps 60 --e2e-steps 0 2> gpurun_out/b.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c5 twolayer', {k:d[k] for k in ('ms_per_step','kernel_ms_per_step','punctured_simplices','cells_refined_per_step')}, d['roofline']['frac'])"
