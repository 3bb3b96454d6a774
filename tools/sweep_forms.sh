#!/bin/bash
for form in 1 2; do
  for prec in bf16x3 bf16; do
    VTC_B200_FORMULATION=$form timeout 300 python bench.py --steps 3 --warmup 2 --no-extras --precision $prec 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('form=$form', '$prec', 'ms_per_step=%.1f patches/s=%.0f launch_ms=%.4f' % (d['ms_per_step'], d['value'], d['roofline']['launch_ms']))"
  done
done
