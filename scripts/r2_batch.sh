#!/bin/bash
# Gathers of a whole quadrature row in flight (wide path): default (8 rows x 4 CTAs, 64 registers: ptxas interleaves two samples, 1-2 gathers in
# flight per warp) vs 2 CTAs at 128 registers (ptxas hoists the five gathers of a row by itself: nb_m2; the same with the explicit batch:
# wb_m2) vs 3 CTAs at 80 registers with the explicit batch (3+2: wb_m3)
mkdir -p gpurun_out
: > gpurun_out/r2_batch_ab.txt
for c in "2160 3840 3 5 300" "480 640 3 5 0" "480 640 3 5 6000"; do
  python scripts/ab3.py t8 $c >> gpurun_out/r2_batch_ab.txt 2>&1
  for v in nb_m2 wb_m2 wb_m3; do QGMAP_LIB_PATH=build/libqgmap_$v.so python scripts/ab3.py $v $c >> gpurun_out/r2_batch_ab.txt 2>&1; done
done
cat gpurun_out/r2_batch_ab.txt
