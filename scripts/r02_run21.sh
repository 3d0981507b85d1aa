#!/bin/bash
# round 2, GPU session 21 (1 GPU): key-filtered min non-zero |v| in the scalar build kernels (FTKB_RES_FILTER=0: threshold +inf),
# parity on the new kernels, streaming with keys-only staging
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r02u_pytest_gpu.log
show() { python - "$1" "$2" <<'P'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], "ms/step %.4f scan %.4f frac %.3f value %.3e" % (d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["value"]))
P
}
for f in 1 0; do
  FTKB_RES_FILTER=$f timeout 200 python bench.py --config c2 --only-main --steps 60 --warmup 5 --e2e-steps 0 2>/dev/null | tail -1 > gpurun_out/r02u_c2_filter$f.json
  show gpurun_out/r02u_c2_filter$f.json "c2 filter=$f"
  FTKB_RES_FILTER=$f timeout 200 python bench.py --config c3 --only-main --steps 24 --warmup 4 --e2e-steps 0 2>/dev/null | tail -1 > gpurun_out/r02u_c3_filter$f.json
  show gpurun_out/r02u_c3_filter$f.json "c3 filter=$f"
  FTKB_RES_FILTER=$f timeout 300 python bench.py --config woven --only-main --steps 12 --warmup 3 --e2e-steps 0 2>/dev/null | tail -1 > gpurun_out/r02u_woven_filter$f.json
  show gpurun_out/r02u_woven_filter$f.json "woven filter=$f"
done
timeout 300 python scripts/stream_timing.py 4096 4096 32 > gpurun_out/r02u_stream_timing.jsonl 2>gpurun_out/r02u_stream.err
python - <<'P'
import json
for l in open("gpurun_out/r02u_stream_timing.jsonl"):
    d = json.loads(l)
    print(d["mode"], "ms/timestep %.3f host %.2f dev %.2f scan %.2f pts %d traj %d d2h %d" % (d["ms_per_timestep"], d["ms_host_trace"], d["ms_device_trace"], d["ms_scan"], d["punctured"], d["trajectories"], d["d2h_bytes"]))
P
