#!/bin/bash
# round 2, GPU session 29 (1 GPU): compute-sanitizer on the final ring kernels (key filter with shared-memory accumulators, survivor staging)
mkdir -p gpurun_out
for tool in racecheck synccheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_rings.py > gpurun_out/r02g_sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; tail -4 gpurun_out/r02g_sanitizer_$tool.log
  grep -c "Race reported\|ERROR SUMMARY" gpurun_out/r02g_sanitizer_$tool.log
done
grep -h "Race reported" gpurun_out/r02g_sanitizer_racecheck.log | sed 's/.*Race reported between //' | sed 's/ at .* and / vs /' | sed 's/ at .*//' | sort | uniq -c | sort -rn | head
grep -h -A3 "Race reported" gpurun_out/r02g_sanitizer_racecheck.log | grep "at \|in " | sed 's/+0x.*//' | sort | uniq -c | sort -rn | head -12
