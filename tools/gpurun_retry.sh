#!/bin/bash
# tools/gpurun_retry.sh LOG [gpurun args...]: retry a gpurun call while the pod answers busy / transient (rc 3)
log=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient\|rc=3\|busy" "$log" && ! grep -q "status=ok\|status=fail" "$log"; then sleep 90; continue; fi
  break
done
echo "done rc=$rc" >> "$log"
