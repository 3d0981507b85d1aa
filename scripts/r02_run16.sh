#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_group.py -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r02p_pytest_group.log
