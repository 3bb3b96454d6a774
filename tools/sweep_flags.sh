#!/bin/bash
# usage: tools/sweep_flags.sh "0 1 2 4 7" -> per-launch ms of the fused FISTA kernel for each VTC_B200_FLAGS value
for f in $1; do
  for prec in bf16x3 bf16; do
    VTC_B200_FLAGS=$f timeout 300 python bench.py --steps 2 --warmup 2 --no-extras --precision $prec 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('flags=$f', '$prec', 'ms_per_step=%.1f launch_ms=%.4f' % (d['ms_per_step'], d['roofline']['launch_ms']))"
  done
done
