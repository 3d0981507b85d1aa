#!/bin/bash
# round 2, GPU session 55 (1 GPU): streaming timing and the CLI raw-input timing on the last library
mkdir -p gpurun_out
timeout 300 python scripts/stream_timing.py 4096 4096 32 > gpurun_out/r02af2_stream_timing.jsonl 2>gpurun_out/r02af2_stream.err
python - <<'P'
import json
for l in open("gpurun_out/r02af2_stream_timing.jsonl"):
    d = json.loads(l)
    print(d["mode"], "ms/timestep %.3f host %.2f dev %.2f scan %.2f test %.2f pts %d traj %d" % (d["ms_per_timestep"], d["ms_host_trace"], d["ms_device_trace"], d["ms_scan"], d["ms_test"], d["punctured"], d["trajectories"]))
P
timeout 300 python -m pytest tests/test_streaming.py tests/test_cli.py -m gpu -q -x 2>&1 | tail -2
TAG=r02af2 timeout 400 bash scripts/cli_input_timing.sh 2>&1 | tail -8
