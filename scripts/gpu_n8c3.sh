#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --config c3 --steps 31 --warmup 3 --e2e-steps 4 2> gpurun_out/bench_c3_n8.err | tee gpurun_out/bench_c3_n8_peer.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c3 n8 peer', d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['trajectories'], d['punctured_simplices'], d['e2e']['value'], d['config']['halo'][:20])"
grep -i "error\|Traceback" gpurun_out/bench_c3_n8.err | head -3
