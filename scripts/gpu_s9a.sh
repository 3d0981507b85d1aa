#!/bin/bash
# session 9: parity of the range-cell 3D scan + A/B against the two-layer fused scan
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "3d or 3D or abc or extremum" 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_3d.log
timeout 300 python bench.py --config c3 --steps 31 --no-cpu-baseline --e2e-steps 4 2> gpurun_out/bench_c3.err | tee gpurun_out/bench_c3_cells.json
tail -3 gpurun_out/bench_c3.err
FTKB_SCAN3D=twolayer timeout 300 python bench.py --config c3 --steps 31 --no-cpu-baseline --e2e-steps 4 2> gpurun_out/bench_c3b.err | tee gpurun_out/bench_c3_twolayer.json
tail -3 gpurun_out/bench_c3b.err
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
