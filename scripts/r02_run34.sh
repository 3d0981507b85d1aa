#!/bin/bash
# round 2, GPU session 34 (1 GPU): sorted records stay on the device until somebody asks for them; full parity tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r02k_pytest_gpu.log
sed -i 's/r02j_/r02k_/g' scripts/r02_run33.sh
bash scripts/r02_run33.sh
