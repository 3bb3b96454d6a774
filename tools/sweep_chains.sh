#!/bin/bash
for ch in 1 2; do
  for prec in bf16x3 bf16; do
    VTC_B200_CHAINS=$ch timeout 300 python bench.py --steps 3 --warmup 2 --no-extras --precision $prec 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chains=$ch', '$prec', 'ms_per_step=%.1f patches/s=%.0f ms_per_iteration=%.4f' % (d['ms_per_step'], d['value'], d['roofline']['ms_per_iteration']))"
  done
done
