#!/bin/bash
# usage: tools/sweep_flags2.sh "0 64" -> ms per step / per iteration / first & fused launch for each VTC_B200_FLAGS value
for rep in 1 2; do
for f in $1; do
  for prec in bf16x3 bf16; do
    VTC_B200_FLAGS=$f timeout 300 python bench.py --steps 3 --warmup 2 --no-extras --precision $prec 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('flags=$f', '$prec', 'ms_per_step=%.1f iter=%.4f first=%.4f fused=%.4f' % (d['ms_per_step'], r['ms_per_iteration'], r['first_launch_ms'] or 0, r['launch_ms']))"
  done
done
done
