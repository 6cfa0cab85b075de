#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3: nothing charged). usage: tools/gpu_retry.sh [gpurun args...]
make -s -j8 -C "$(dirname "$0")/../spade_b200/csrc" > /dev/null || { echo 'gpu_retry: library build failed'; exit 1; }
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@"; rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
