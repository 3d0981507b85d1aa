#!/bin/bash
# N = 2 check of the default multi-GPU bench path (halo through peer memory) and of the NCCL copy
mkdir -p gpurun_out
for halo in peer nccl; do
FTKB_HALO=$halo timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 126 --warmup 3 2> gpurun_out/bench_n2.err | tee gpurun_out/bench_c2_n2_$halo.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c2 n2 $halo', d['value'], d['ms_per_step'], d['trajectories'], d['punctured_simplices'], d['config']['halo'][:24])"
grep -i "error\|Traceback" gpurun_out/bench_n2.err | head -3
done
