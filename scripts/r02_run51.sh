#!/bin/bash
# round 2, GPU session 51 (1 GPU): finalize phase laps inside the default bench line (the dense record's finalize is 10x its stand-alone time there)
mkdir -p gpurun_out
FTKB_DEBUG_TIMING=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 0 2> gpurun_out/r02ab2_bench.err | cut -c1-200
grep "ftkb timing" gpurun_out/r02ab2_bench.err | tail -12
