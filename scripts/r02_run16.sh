#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_group.py -m gpu -q 2>&1 | tail -15 | tee gpurun_out/r02p_pytest_group.log
# e2e of one tracker over z-slabs on this box's GPUs (device list repeated on one GPU: functional only)
