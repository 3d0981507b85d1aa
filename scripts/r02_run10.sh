#!/bin/bash
# round 2, GPU session 10 (2 GPUs): the driver's launch of bench.py at N=2 (peer-memory halo, sub-records, self-check), group / distributed tests
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/r02j_gpus.log
timeout 900 python -m pytest tests/test_group.py tests/test_distributed.py -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/r02j_pytest_gpu_2gpus.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 \
    2> gpurun_out/r02j_bench_n2.err | tee gpurun_out/r02j_bench_n2.json | cut -c1-300
tail -5 gpurun_out/r02j_bench_n2.err
timeout 900 python bench.py --steps 20 --warmup 5 2> gpurun_out/r02j_bench_n1.err | tee gpurun_out/r02j_bench_n1.json | cut -c1-300
bash scripts/cli_input_timing.sh gpurun_out
ls -la gpurun_out | tail -3
