#!/bin/bash
# memcheck of the scan / build / test kernels on small parity cases (slow: a few minutes)
mkdir -p gpurun_out
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "range_cell_scan or direct_staging or matches_oracle_random or non_finite or single_snapshot or buffers_grow" > gpurun_out/memcheck.log 2>&1
echo "memcheck rc=$?"
grep -c "Invalid\|out of bounds" gpurun_out/memcheck.log; grep "ERROR SUMMARY\|passed\|failed" gpurun_out/memcheck.log | tail -5
grep -B2 -A12 "Invalid" gpurun_out/memcheck.log | head -60
