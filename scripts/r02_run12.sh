#!/bin/bash
# round 2, GPU session 12 (8 GPUs): the driver's launch of bench.py at N=8 and N=4
mkdir -p gpurun_out
nvidia-smi -L | wc -l | tee gpurun_out/r02l_gpus.log
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 \
    2> gpurun_out/r02l_bench_n$n.err | tee gpurun_out/r02l_bench_n$n.json | cut -c1-250
tail -3 gpurun_out/r02l_bench_n$n.err
done
timeout 600 python -m pytest tests/test_group.py -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/r02l_pytest_group_8gpus.log
ls -la gpurun_out | tail -3
