#!/bin/bash
# K >= 7 tile without a halo warp: 8 rows x 3 CTAs at 80 registers (default) vs 9 rows x 3 CTAs at 72 registers (27 warps per SM)
mkdir -p gpurun_out
C="full:480:640:2:9:0:g,full:480:640:2:9:4000:g,full:388:584:2:9:4000:g,full:480:640:2:7:3000:g,full:480:640:3:11:3000:g"
: > gpurun_out/r2_bigk_ab.txt
for rep in 1 2; do
python scripts/ab2.py t8 "$C" "t8=" >> gpurun_out/r2_bigk_ab.txt 2>&1
QGMAP_LIB_PATH=build/libqgmap_bk9.so python scripts/ab2.py t9 "$C" "t9=" >> gpurun_out/r2_bigk_ab.txt 2>&1
done
sort -k3,8 -s gpurun_out/r2_bigk_ab.txt
