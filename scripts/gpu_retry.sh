#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3: nothing charged).
# usage: scripts/gpu_retry.sh <timeout_s> [--gpus N] -- '<command>'
T=$1; shift
for i in $(seq 1 40); do
    /usr/local/graft/bin/gpurun --timeout "$T" "$@"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    echo "[gpu_retry] busy, attempt $i; sleeping 45 s" >&2
    sleep 45
done
exit 3
