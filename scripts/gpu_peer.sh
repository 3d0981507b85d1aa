#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "peer_memory or time_slab" 2>&1 | tail -8
for cfg in c2 c4; do
extra=""; steps=126; if [ $cfg = c4 ]; then extra="--e2e-steps 0"; steps=8; fi
for halo in peer nccl; do
FTKB_HALO=$halo timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --config $cfg --steps $steps --warmup 3 $extra 2> gpurun_out/bench_peer.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$cfg n2 halo=$halo', d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['trajectories'], d['punctured_simplices'], d['config'].get('halo','')[:30])"
grep -i "error\|Traceback" gpurun_out/bench_peer.err | head -3
done
done
