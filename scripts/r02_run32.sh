#!/bin/bash
# round 2, GPU session 32 (8 GPUs): the driver's launch of bench.py at N = 8 on the final library
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 \
    2> gpurun_out/r02i_bench_n8.err | tee gpurun_out/r02i_bench_n8.json | cut -c1-300
tail -3 gpurun_out/r02i_bench_n8.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/r02i_bench_n8.json").read().strip().splitlines()[-1])
print("n8 c2", d["value"], d["ms_per_step"], "selfcheck", d.get("selfcheck"), "e2e", d.get("e2e"))
for k in ("scaling_c4", "scaling_c3", "dense_woven"):
    r = d.get(k) or {}
    print(k, r.get("value"), r.get("ms_per_step"), r.get("trajectories"), r.get("punctured_simplices"))
P
