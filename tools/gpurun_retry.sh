#!/usr/bin/env bash
# tools/gpurun_retry.sh [--timeout S] -- '<command>': gpurun, retried while the pod answers "transient" / busy
# (nothing is charged for those answers).  Gives up after 12 tries.
for try in $(seq 1 12); do
  out="$(/usr/local/graft/bin/gpurun "$@" 2>&1)"; rc=$?
  if echo "$out" | grep -q "status=transient\|status=busy" || [ $rc -eq 3 ]; then sleep 150; continue; fi
  echo "$out"; exit $rc
done
echo "$out"; exit 3
