#!/bin/bash
# round 2, GPU session 27 (1 GPU): origin-in-simplex through the unsorted determinants (full cascade only for zero / -2^63 determinants)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r02aa_pytest_gpu.log
show() { python - "$1" "$2" <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms/step %.4f scan %.4f frac %.3f value %.3e" % (d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["value"]), d.get("kernel_ms_per_step"), "repeated", d["roofline"].get("sweeps_repeated"), d.get("trajectories"), d.get("punctured_simplices"))
except Exception as e:
    print(sys.argv[2], "failed", e)
P
}
for k in 1 2; do
  timeout 300 python bench.py --config woven --only-main --steps 12 --warmup 3 --e2e-steps 0 2>/dev/null | tail -1 > gpurun_out/r02aa_woven_$k.json
  show gpurun_out/r02aa_woven_$k.json "woven run $k"
done
timeout 200 python bench.py --config c2 --only-main --steps 60 --warmup 5 --e2e-steps 0 2>/dev/null | tail -1 > gpurun_out/r02aa_c2.json
show gpurun_out/r02aa_c2.json "c2"
timeout 200 python bench.py --config c3 --only-main --steps 24 --warmup 4 --e2e-steps 0 2>/dev/null | tail -1 > gpurun_out/r02aa_c3.json
show gpurun_out/r02aa_c3.json "c3"
timeout 300 python scripts/stream_timing.py 4096 4096 32 > gpurun_out/r02aa_stream_timing.jsonl 2>gpurun_out/r02aa_stream.err
python - <<'P'
import json
for l in open("gpurun_out/r02aa_stream_timing.jsonl"):
    d = json.loads(l)
    print(d["mode"], "ms/timestep %.3f host %.2f dev %.2f scan %.2f test %.2f pts %d traj %d" % (d["ms_per_timestep"], d["ms_host_trace"], d["ms_device_trace"], d["ms_scan"], d["ms_test"], d["punctured"], d["trajectories"]))
P
