#!/bin/bash
mkdir -p gpurun_out
# second decode: launches 24.. ; k7 C=512 = 24+9, k7 C=256 = 24+16, k1 C=256 = 24+17, convT 128->128 = 24+23
for spec in "33 k7_c512" "40 k7_c256" "41 k1_c256" "47 convT_c128_s2"; do
  set -- $spec
  timeout 280 ncu --set full --import-source on --clock-control none -k regex:conv_umma2 -s $1 -c 1 -f -o gpurun_out/prof_umma2_$2 python tools/prof_decode.py 8 > gpurun_out/ncu_umma2_$2.log 2>&1; echo "$2 exit $?"
done
