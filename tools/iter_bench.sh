#!/bin/bash
# usage: tools/iter_bench.sh tag   -- parity spot checks + bench of the one-launch schedule (and the two-launch one with ALSO_OLD=1)
tag=$1
mkdir -p gpurun_out
for c in "512 1024 256 20" "300 1000 250 20 bf16x3 ista" "1000 96 20 20" "4096 1024 256 50 bf16"; do
  timeout 300 python tools/iter_debug.py $c 2>&1 | tail -3
done
for f in ${ALSO_OLD:+0} 1; do for p in bf16x3 bf16; do
  VTC_B200_FUSED_ITER=$f timeout 600 python bench.py --no-extras --precision $p --steps 3 --warmup 3 > gpurun_out/bench_${tag}_f${f}_$p.json 2> gpurun_out/bench_${tag}_f${f}_$p.err
  python - <<PY
import json
try:
  d=json.load(open("gpurun_out/bench_${tag}_f${f}_$p.json")); r=d["roofline"]
  print("fused=$f $p ms/step %.2f iter %.4f launch %.4f first %s frac %.3f %s clocks %s" % (d["ms_per_step"], r["ms_per_iteration"], r["launch_ms"], r["first_launch_ms"], r["frac"], r["bound"], d["clocks"]["sm_mhz"]))
except Exception as e:
  print("fused=$f $p failed", e); print(open("gpurun_out/bench_${tag}_f${f}_$p.err").read()[-1500:])
PY
done; done
