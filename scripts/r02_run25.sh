#!/bin/bash
# round 2, GPU session 25 (1 GPU): 2D cold path with every load of a layer issued before the first use
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_wrap.py tests/test_pyftk.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r02y_pytest_gpu.log
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
  timeout 300 python bench.py --config woven --only-main --steps 12 --warmup 3 --e2e-steps 0 2>/dev/null | tail -1 > gpurun_out/r02y_woven_$k.json
  show gpurun_out/r02y_woven_$k.json "woven run $k"
done
FTKB_TEST_OVERLAP=0 timeout 300 python bench.py --config woven --only-main --steps 12 --warmup 3 --e2e-steps 0 2>/dev/null | tail -1 > gpurun_out/r02y_woven_nooverlap.json
show gpurun_out/r02y_woven_nooverlap.json "woven, one stream"
timeout 300 python scripts/stream_timing.py 4096 4096 32 > gpurun_out/r02y_stream_timing.jsonl 2>gpurun_out/r02y_stream.err
python - <<'P'
import json
for l in open("gpurun_out/r02y_stream_timing.jsonl"):
    d = json.loads(l)
    print(d["mode"], "ms/timestep %.3f host %.2f dev %.2f scan %.2f test %.2f pts %d traj %d" % (d["ms_per_timestep"], d["ms_host_trace"], d["ms_device_trace"], d["ms_scan"], d["ms_test"], d["punctured"], d["trajectories"]))
P
