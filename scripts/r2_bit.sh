#!/bin/bash
mkdir -p gpurun_out
python scripts/bitcmp.py build/libqgmap_halo1.so gqmap-opticalflow_b200/libqgmap.so > gpurun_out/r2_bitcmp.txt 2>&1; cat gpurun_out/r2_bitcmp.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_nh2_pytest.log 2>&1; tail -5 gpurun_out/r2_nh2_pytest.log
C="full:2160:3840:3:5:300:g,full:480:640:3:5:0:g,full:480:640:3:5:6000:g,full:388:584:1:3:1000:g,full:480:640:2:9:4000:g,full:480:640:2:9:0:g,full:480:640:2:7:3000:g,full:480:640:2:4:3000:g,full:2160:3840:3:5:300,super:480:640:3:5:3000:g"
: > gpurun_out/r2_nh2_ab.txt
python scripts/ab2.py new "$C" "new=" >> gpurun_out/r2_nh2_ab.txt 2>&1
QGMAP_LIB_PATH=build/libqgmap_halo1.so python scripts/ab2.py halo1 "$C" "halo1=" >> gpurun_out/r2_nh2_ab.txt 2>&1
sort -k3,8 -s gpurun_out/r2_nh2_ab.txt
