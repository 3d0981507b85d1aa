#!/bin/bash
# round 2, GPU session 46 (1 GPU): finalize scratch planned in one allocation; full parity tests; finalize laps on the dense field
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r02v2_pytest_gpu.log
sed -i 's/r02u2_/r02v2_/g' scripts/r02_run45.sh
bash scripts/r02_run45.sh
timeout 300 python bench.py --config c5 --only-main --steps 60 --warmup 3 --e2e-steps 0 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c5', d['ms_per_step'], {k: d.get(k) for k in ('finalize_ms','finalize_ms_library','finalize_ms_device','finalize_ms_host','punctured_simplices')})"
