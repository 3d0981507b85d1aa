#!/bin/bash
mkdir -p gpurun_out
FTKB_HALO=peer timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --config c4 --steps 8 --warmup 3 --e2e-steps 0 2> gpurun_out/bench_peer8.err | tee gpurun_out/bench_c4_n8_peer.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c4 n8 peer', d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['trajectories'], d['punctured_simplices'])"
grep -i "error\|Traceback" gpurun_out/bench_peer8.err | head -3
FTKB_HALO=peer timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 252 --warmup 3 2> gpurun_out/bench_peer8b.err | tee gpurun_out/bench_c2_n8_peer.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c2 n8 peer', d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['trajectories'], d['punctured_simplices'], d['e2e']['value'])"
grep -i "error\|Traceback" gpurun_out/bench_peer8b.err | head -3
