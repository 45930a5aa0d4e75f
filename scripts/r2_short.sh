#!/bin/bash
# K=5 tile height without a halo warp, shorter tiles: 8 rows x 4 CTAs (default) vs 6 rows x 5 CTAs vs 4 rows x 8 CTAs (all 64 registers)
mkdir -p gpurun_out
C="full:2160:3840:3:5:300:g,full:480:640:3:5:0:g,full:480:640:3:5:6000:g"
: > gpurun_out/r2_short_ab.txt
python scripts/ab2.py t8 "$C" "t8=" >> gpurun_out/r2_short_ab.txt 2>&1
for v in k5t6 k5t4; do QGMAP_LIB_PATH=build/libqgmap_$v.so python scripts/ab2.py $v "$C" "$v=" >> gpurun_out/r2_short_ab.txt 2>&1; done
sort -k3,8 -s gpurun_out/r2_short_ab.txt
