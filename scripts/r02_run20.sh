#!/bin/bash
# round 2, GPU session 20 (1 GPU): streaming grow step with the batch prepared on the device (sort + neighbour lists) against
# the host-hash preparation; streaming / parity tests on the new path
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_streaming.py tests/test_coords.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r02t_pytest_streaming.log
FTKB_STREAM_PREP=host timeout 300 python -m pytest tests/test_streaming.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python scripts/stream_timing.py 4096 4096 32 > gpurun_out/r02t_stream_timing_device_prep.jsonl 2>gpurun_out/r02t_stream.err
FTKB_STREAM_PREP=host timeout 300 python scripts/stream_timing.py 4096 4096 32 > gpurun_out/r02t_stream_timing_host_prep.jsonl 2>>gpurun_out/r02t_stream.err
FTKB_STREAM_GROW=async timeout 300 python scripts/stream_timing.py 4096 4096 32 > gpurun_out/r02t_stream_timing_device_prep_async.jsonl 2>>gpurun_out/r02t_stream.err
python - <<'P'
import json
for f in ("device_prep", "host_prep", "device_prep_async"):
    for l in open(f"gpurun_out/r02t_stream_timing_{f}.jsonl"):
        d = json.loads(l)
        print(f, d["mode"], "ms/timestep %.3f host %.2f dev %.2f scan %.2f pts %d traj %d" % (d["ms_per_timestep"], d["ms_host_trace"], d["ms_device_trace"], d["ms_scan"], d["punctured"], d["trajectories"]))
P
tail -5 gpurun_out/r02t_stream.err
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r02t_pytest_gpu.log
