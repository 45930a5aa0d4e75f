#!/bin/bash
# lighter block-finish ticket (release atomic after a warp barrier, no MEMBAR.SC, no L1 invalidate per CTA): parity + A/B against the previous build
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bands.py tests/test_gpu_edge_cases.py tests/test_gpu_walk.py -x -q -m gpu > gpurun_out/r2_fin_pytest.log 2>&1; tail -3 gpurun_out/r2_fin_pytest.log
C="full:480:640:3:5:6000:g,full:480:640:3:5:0:g,full:388:584:1:3:1000:g,full:480:640:2:9:4000:g,super:480:640:3:5:3000:g,full:2160:3840:3:5:300:g"
: > gpurun_out/r2_fin_ab.txt
for rep in 1 2; do
python scripts/ab2.py light "$C" "light=" >> gpurun_out/r2_fin_ab.txt 2>&1
QGMAP_LIB_PATH=build/libqgmap_nopipe.so python scripts/ab2.py old "$C" "old=" >> gpurun_out/r2_fin_ab.txt 2>&1
done
sort -k3,8 -s gpurun_out/r2_fin_ab.txt
