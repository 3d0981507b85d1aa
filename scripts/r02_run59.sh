#!/bin/bash
# round 2, last GPU session: the whole GPU suite and smoke on the final commit
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/r02final_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/r02final_smoke.log
