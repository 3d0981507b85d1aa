#!/bin/bash
# round 2, GPU session 52 (1 GPU): the default bench line with finalize phase laps (where do the dense record's 196 ms go?)
mkdir -p gpurun_out
FTKB_DEBUG_TIMING=1 timeout 900 python bench.py 2> gpurun_out/r02ac2_bench.err | cut -c1-200
grep "ftkb timing" gpurun_out/r02ac2_bench.err | tail -12
