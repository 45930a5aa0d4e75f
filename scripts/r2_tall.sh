#!/bin/bash
# K=5 tile height without a halo warp: 8 rows x 4 CTAs (default) vs 16 rows x 2 CTAs (64 regs) vs 12 rows x 2 CTAs (80 regs, fewer spills)
mkdir -p gpurun_out
C="full:2160:3840:3:5:300:g,full:480:640:3:5:0:g,full:480:640:3:5:6000:g,full:1080:1920:3:5:300:g"
: > gpurun_out/r2_tall_ab.txt
python scripts/ab2.py t8 "$C" "t8=" >> gpurun_out/r2_tall_ab.txt 2>&1
for v in k5t16 k5t12; do QGMAP_LIB_PATH=build/libqgmap_$v.so python scripts/ab2.py $v "$C" "$v=" >> gpurun_out/r2_tall_ab.txt 2>&1; done
sort -k3,8 -s gpurun_out/r2_tall_ab.txt
