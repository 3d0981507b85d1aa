#!/bin/bash
# round 2, GPU session 17 (8 GPUs): z-slab / time-chunk groups on real device lists, two-process peer-memory tests on two GPUs, group bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_group.py tests/test_distributed.py -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r02q_pytest_multi_gpu.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "peer_memory or time_slab" 2>&1 | tail -4 | tee -a gpurun_out/r02q_pytest_multi_gpu.log
timeout 420 python scripts/group_bench.py --steps 12 2> gpurun_out/r02q_group_bench.err | tee gpurun_out/r02q_group_bench.jsonl | cut -c1-330
tail -3 gpurun_out/r02q_group_bench.err
