#!/bin/bash
# round 2, GPU session 54 (2 GPUs, last library): multi-GPU paths on the final library -- two-process tests (peer-memory halo, NCCL), groups behind the
# C ABI, and the driver's launch of bench.py at N = 2
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_distributed.py tests/test_group.py tests/test_gpu_parity.py -m gpu -q -x -k "slab or peer or group or shard or distributed or halo or device" 2>&1 | tail -4 | tee gpurun_out/r02ae2_pytest_multi_gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 \
    2> gpurun_out/r02ae2_bench_n2.err | tee gpurun_out/r02ae2_bench_n2.json | cut -c1-300
tail -3 gpurun_out/r02ae2_bench_n2.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/r02ae2_bench_n2.json").read().strip().splitlines()[-1])
print("n2 c2", d["value"], d["ms_per_step"], "selfcheck", d.get("selfcheck"))
for k in ("scaling_c4", "scaling_c3", "dense_woven"):
    r = d.get(k) or {}
    print(k, r.get("value"), r.get("ms_per_step"), r.get("trajectories"), r.get("punctured_simplices"))
P
