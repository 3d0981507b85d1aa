#!/bin/bash
mkdir -p gpurun_out
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 252 --warmup 3 2> gpurun_out/bench_n$n.err | tee gpurun_out/bench_c2_n$n.json | cut -c1-300
grep -v "Warning\|OMP_NUM\|\*\*\*\*" gpurun_out/bench_n$n.err | tail -3
done
