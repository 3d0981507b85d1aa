#!/bin/bash
# round 2, GPU session 57 (1 GPU): the feature-dense parity test (white noise: staged survivors, dense test kernels in line, buffer growth ahead)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "feature_dense or overflow" 2>&1 | tail -4 | tee gpurun_out/r02ah2_pytest_dense.log
FTKB_DEFER=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "feature_dense" 2>&1 | tail -2
