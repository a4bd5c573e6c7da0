#!/bin/bash
# retries a gpurun call while the pod answers "busy" (exit code 3); usage: gpurun_retry.sh [gpurun args] -- 'command'
for i in $(seq 1 20); do
    /usr/local/graft/bin/gpurun "$@"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    sleep 60
done
exit 3
