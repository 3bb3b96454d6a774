#!/bin/bash
# one process per case: a trapped kernel poisons the CUDA context
mkdir -p gpurun_out
for c in "256 128 64 1" "256 128 64 2" "256 128 64 5" "256 256 256 10" "512 1024 256 20" "300 1000 250 20 bf16x3 ista" \
         "1000 96 20 20" "512 1024 256 20 bf16" "65536 1024 256 100" "65536 1024 256 100 bf16"; do
  timeout 300 python tools/iter_debug.py $c 2>&1 | tail -4
done
