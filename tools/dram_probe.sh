#!/bin/bash
# usage: tools/dram_probe.sh "0 8 16" -> DRAM bytes of one steady-state fused FISTA launch for each VTC_B200_FLAGS value
for f in $1; do
  for prec in bf16x3 bf16; do
    VTC_B200_FLAGS=$f ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:vtc_gemm_kernel -s 7 -c 1 --csv python tools/profile_fista.py --precision $prec 2>/dev/null | grep -E "dram__|gpu__time" | awk -F'","' -v f=$f -v p=$prec '{gsub(/"/,"",$NF); printf "flags=%s %s %s %s %s\n", f, p, $(NF-2), $(NF-1), $NF}'
  done
done
