#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 126 --warmup 3 2> gpurun_out/bench_n2.err | tee gpurun_out/bench_c2_n2.json
tail -5 gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config c3 --steps 31 --warmup 3 2> gpurun_out/bench_n2c3.err | tee gpurun_out/bench_c3_n2.json
tail -5 gpurun_out/bench_n2c3.err
timeout 600 python -m pytest tests -m gpu -q -x -k "shard or distrib or slab" 2>&1 | tail -4
