#!/bin/bash
# what each piece of the persistent iteration kernel costs in situ: time with the piece left out (results are wrong)
#   1 state loads  2 state stores  4 r_k re-streaming  8 state LDS/STS  16 R MMAs  32 G MMAs  64 y STS  128 tcgen05.ld
for a in ${ABLATES:-0 1 2 3 4 8 11 15 7 16 32 48 64 128 79 207 255}; do
  VTC_B200_ABLATE=$a timeout 300 python tools/ablate.py ${BATCH:-65536} ${ITERS:-60} 2>&1 | tail -1
done
