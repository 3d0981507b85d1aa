#!/bin/bash
# round 2, GPU session 56 (1 GPU): the default bench line and the reference arm after the CPU-baseline sample change
mkdir -p gpurun_out
( time timeout 900 python bench.py 2> gpurun_out/r02ag2_bench_c2.err > gpurun_out/r02ag2_bench_c2.json ) 2>&1 | grep real
python - <<'P'
import json
d = json.loads(open("gpurun_out/r02ag2_bench_c2.json").read().strip().splitlines()[-1])
print("c2", d["ms_per_step"], d["roofline"]["frac"], d["cpu_baseline"], d["e2e"]["ms_per_step"])
for k in ("scaling_c4", "scaling_c3", "dense_woven"):
    r = d[k]; print(k, r["value"], r["ms_per_step"], r.get("finalize_ms"))
P
tail -2 gpurun_out/r02ag2_bench_c2.err
( time timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2> gpurun_out/r02ag2_bench_ref.err | tee gpurun_out/r02ag2_bench_reference_arm.json | cut -c1-300 ) 2>&1 | tail -4
