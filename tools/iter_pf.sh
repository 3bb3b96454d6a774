#!/bin/bash
for pf in ${PFS:-0 8 16 32 64}; do
  echo "=== prefetch distance $pf"
  VTC_B200_ITER_PREFETCH=$pf timeout 200 python tools/iter_debug.py 600 1000 250 25 2>&1 | tail -2
  for v in ${VARIANTS:-0 1}; do
    echo "variant $v"; VTC_B200_ITER_PREFETCH=$pf VTC_B200_ITER_VARIANT=$v timeout 300 python tools/iter_times.py 65536 2>&1 | grep "fused=1"
  done
done
