#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --no-cpu-baseline --e2e-steps 2 2> gpurun_out/b.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c2', {k:d[k] for k in ('ms_per_step','kernel_ms_per_step')}, d['roofline']['frac'])"; tail -2 gpurun_out/b.err
timeout 300 python bench.py --config c4s --steps 20 --e2e-steps 2 2> gpurun_out/b.err | tee gpurun_out/bench_c4s.json | cut -c1-400; tail -2 gpurun_out/b.err
timeout 600 python bench.py --config c5 --steps 60 --e2e-steps 4 2> gpurun_out/b.err | tee gpurun_out/bench_c5.json | cut -c1-1500; tail -2 gpurun_out/b.err
timeout 900 python bench.py --config c4 --steps 12 --e2e-steps 0 2> gpurun_out/b.err | tee gpurun_out/bench_c4.json | cut -c1-1500; tail -2 gpurun_out/b.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
