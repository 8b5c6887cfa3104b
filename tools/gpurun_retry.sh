#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3 / "transient").  usage: tools/gpurun_retry.sh <log> <gpurun args...>
# The repo is snapshotted at each attempt: only start this from a consistent tree, and touch /tmp/b2d_hold to pause retries.
LOG=$1; shift
for i in $(seq 1 30); do
  while [ -e /tmp/b2d_hold ]; do sleep 10; done
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1; rc=$?
  if grep -q "status=transient" "$LOG" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  if [ $rc -eq 2 ] && grep -qi "another call\|in flight\|running" "$LOG"; then sleep 60; continue; fi
  echo "gpurun finished rc=$rc (attempt $i)" >> "$LOG"; exit $rc
done
echo "gave up after 30 attempts" >> "$LOG"
