#!/bin/bash
# what binds the wide regime at 4K: tile shapes (L1 footprint per SM) A/B, and a --set full capture with the whole raw metric page kept
mkdir -p gpurun_out
C="full:2160:3840:3:5:300:g,full:480:640:3:5:0:g,full:480:640:3:5:6000:g"
: > gpurun_out/r2_l2_ab.txt
for v in nopipe t15m2 t31m1; do QGMAP_LIB_PATH=build/libqgmap_$v.so python scripts/ab2.py $v "$C" "$v=" >> gpurun_out/r2_l2_ab.txt 2>&1; done
python scripts/ab2.py pipe "$C" "pipe=" >> gpurun_out/r2_l2_ab.txt 2>&1
cat gpurun_out/r2_l2_ab.txt
R=/tmp/reps; mkdir -p $R
timeout 900 ncu --set full --clock-control none -k regex:qgmap_iter -s 300 -c 1 -o $R/pipe4k -f python scripts/profile_target.py full 3 5 2160 3840 2 300 g > gpurun_out/r2_l2_ncu.log 2>&1
ncu -i $R/pipe4k.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > gpurun_out/r2_l2_pipe4k_raw.csv.gz
python scripts/ncu_summary.py $R/pipe4k.ncu-rep gpurun_out/r2_l2_pipe4k.txt "pipelined wide path, 4K" > /dev/null 2>&1
QGMAP_LIB_PATH=build/libqgmap_t15m2.so timeout 900 ncu --set full --clock-control none -k regex:qgmap_iter -s 300 -c 1 -o $R/t15 -f python scripts/profile_target.py full 3 5 2160 3840 2 300 g >> gpurun_out/r2_l2_ncu.log 2>&1
ncu -i $R/t15.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > gpurun_out/r2_l2_t15_raw.csv.gz
ls -la gpurun_out/r2_l2_*
