#!/bin/bash
# quick GPU check: parity tests + short C2/C3 bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 60 --no-cpu-baseline --e2e-steps 4 2> gpurun_out/bench.err | tee gpurun_out/bench_quick_c2.json | cut -c1-400
tail -3 gpurun_out/bench.err
timeout 300 python bench.py --config c3 --steps 31 --no-cpu-baseline --e2e-steps 4 2> gpurun_out/bench_c3.err | tee gpurun_out/bench_quick_c3.json | cut -c1-400
tail -3 gpurun_out/bench_c3.err
cp MEASURED_PEAKS.json gpurun_out/ 2>/dev/null
nvidia-smi -L
