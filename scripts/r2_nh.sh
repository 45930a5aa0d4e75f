#!/bin/bash
# one-thread-per-belief kernel without a halo warp for K != 3 (single down-edge call site, second trip after the barrier): GPU suite + A/B
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_nh_pytest.log 2>&1; tail -5 gpurun_out/r2_nh_pytest.log
C="full:2160:3840:3:5:300:g,full:480:640:3:5:0:g,full:480:640:3:5:6000:g,full:388:584:1:3:1000:g,full:480:640:2:9:4000:g,full:480:640:2:9:0:g,full:480:640:2:7:3000:g,full:480:640:2:4:3000:g,full:2160:3840:3:5:300"
: > gpurun_out/r2_nh_ab.txt
python scripts/ab2.py new "$C" "new=" >> gpurun_out/r2_nh_ab.txt 2>&1
for v in nh8 g4halo; do QGMAP_LIB_PATH=build/libqgmap_$v.so python scripts/ab2.py $v "$C" "$v=" >> gpurun_out/r2_nh_ab.txt 2>&1; done
sort -k3,8 -s gpurun_out/r2_nh_ab.txt
