#!/bin/bash
# round 2, GPU session 36 (1 GPU): ncu capture of the persistent 2D key kernel (why is it slower than one CTA per chunk?)
mkdir -p gpurun_out
FTKB_K2_PERSIST=1 FTKB_C2_ROWS=27 timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan2d_keys_persist -s 4 -c 1 -o gpurun_out/r02m2_prof_c2persist -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 0 --only-main > gpurun_out/r02m2_ncu_persist.log 2>&1
tail -3 gpurun_out/r02m2_ncu_persist.log
ls -la gpurun_out/r02m2_prof_c2persist.ncu-rep
