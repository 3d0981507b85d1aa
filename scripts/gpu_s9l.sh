#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cli.py tests/test_post_process.py -m gpu -q -x 2>&1 | tail -4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vscan2d_build -s 4 -c 1 -o gpurun_out/prof_v2build -f \
    python bench.py --config c5 --steps 6 --warmup 3 --e2e-steps 0 > gpurun_out/ncu_v2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vscan3d_build -s 4 -c 1 -o gpurun_out/prof_v3build -f \
    python bench.py --config c4 --steps 6 --warmup 3 --e2e-steps 0 > gpurun_out/ncu_v3.log 2>&1
tail -2 gpurun_out/ncu_v3.log
