#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --config c4 --steps 8 --warmup 3 --e2e-steps 0 2> gpurun_out/bench_c4_n8.err | tee gpurun_out/bench_c4_n8.json | cut -c1-400
grep -v "Warning\|OMP_NUM\|\*\*\*\*" gpurun_out/bench_c4_n8.err | tail -5
timeout 600 python bench.py --config c4 --steps 8 --warmup 3 --e2e-steps 0 2> gpurun_out/b.err | tee gpurun_out/bench_c4_n1_k8.json | cut -c1-300
