#!/bin/bash
# builds the library of another commit into tools/ab/libvtc_b200_<name>.so for same-box A/B timing:
#   tools/ab_build.sh <commit> <name>;  VTC_B200_LIB=tools/ab/libvtc_b200_<name>.so python tools/ablate.py
set -e
commit=$1; name=$2
tmp=$(mktemp -d)
git archive "$commit" vision_transform_codes_b200/csrc include | tar -x -C "$tmp"
mkdir -p tools/ab
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC $NVCC_EXTRA \
  -o "tools/ab/libvtc_b200_${name}.so" "$tmp/vision_transform_codes_b200/csrc/vtc_b200.cu"
rm -rf "$tmp"
echo "built tools/ab/libvtc_b200_${name}.so"
