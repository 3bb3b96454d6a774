#!/bin/bash
# parity spot check + per-launch time of every tuning variant of the fused iteration kernel
for v in ${VARIANTS:-0 1 2 3}; do
  echo "=== variant $v"
  VTC_B200_ITER_VARIANT=$v timeout 200 python tools/iter_debug.py 600 1000 250 25 2>&1 | tail -2
  VTC_B200_ITER_VARIANT=$v timeout 300 python tools/iter_times.py ${BATCHES:-65536} 2>&1 | grep "fused=1"
done
