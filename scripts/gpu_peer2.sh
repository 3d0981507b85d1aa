#!/bin/bash
mkdir -p gpurun_out
for cfg in c4 c2; do
extra=""; steps=126; if [ $cfg = c4 ]; then extra="--e2e-steps 0"; steps=8; fi
FTKB_HALO=peer timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --config $cfg --steps $steps --warmup 3 $extra 2> gpurun_out/bench_peer.err | tee gpurun_out/bench_${cfg}_n2_peer.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$cfg n2 halo=peer', d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['trajectories'], d['punctured_simplices'], d['config'].get('halo','')[:30])"
grep -i "error\|Traceback" gpurun_out/bench_peer.err | head -3
done
