#!/bin/bash
mkdir -p gpurun_out
for ce in 0 1; do
NCCL_P2P_USE_CUDA_MEMCPY=$ce timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$ce bench.py --gpus 2 --config c4 --steps 8 --warmup 3 --e2e-steps 0 2> gpurun_out/bench_c4_n2.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c4 n2 CE=$ce', d['value'], d['ms_per_step'], d['kernel_ms_per_step'])"
grep -i "error\|Traceback" gpurun_out/bench_c4_n2.err | head -3
done
