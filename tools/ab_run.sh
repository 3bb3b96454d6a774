#!/bin/bash
# interleaved A/B of library builds / switches on ONE box: each line of CONFIGS is "label|env assignments"
# (run ROUNDS times round-robin so that clock drift hits all alike)
ROUNDS=${ROUNDS:-3}
IFS=$'\n' read -d '' -r -a cfgs <<< "$CONFIGS"
for r in $(seq 1 $ROUNDS); do
  for c in "${cfgs[@]}"; do
    label=${c%%|*}; envs=${c#*|}
    out=$(env $envs timeout 300 python tools/ablate.py ${BATCH:-65536} ${ITERS:-60} 2>&1 | tail -1)
    echo "round $r  $label  ${out#*ms/iter:}"
  done
done
