#!/bin/bash
# usage: scripts/gpu_retry.sh <timeout_s> <logfile> [--gpus N] -- <command>; retries while the pod answers busy (nothing charged)
T=$1; LOG=$2; shift 2
EXTRA=""
if [ "$1" == "--gpus" ]; then EXTRA="--gpus $2"; shift 2; fi
shift   # the "--"
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T $EXTRA -- "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient\|retry in a few minutes" $LOG || [ $rc -eq 3 ]; then sleep 90; continue; fi
  echo "rc=$rc attempt=$i" >> $LOG
  exit $rc
done
echo "gave up" >> $LOG
exit 3
