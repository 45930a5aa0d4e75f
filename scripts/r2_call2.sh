#!/bin/bash
# round 2, GPU call 2: parity suite (row-walking kernel with the halo row inside the walk), register variants, ncu of the 4K problem
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c2_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c2_pytest.log
tail -5 gpurun_out/r2c2_pytest.log
CASES="full:480:640:3:5:2000,full:2160:3840:3:5:300,full:480:640:2:9:4000"
timeout 600 python scripts/ab2.py cur "$CASES" "walk=;tile=QGMAP_ITER:tile" > gpurun_out/r2c2_ab_cur.log 2>&1
for t in w20 w16; do
  QGMAP_LIB_PATH=build/libqgmap_$t.so timeout 600 python scripts/ab2.py $t "$CASES" "walk=;r4=QGMAP_STRIP_ROWS:4;r12=QGMAP_STRIP_ROWS:12" > gpurun_out/r2c2_ab_$t.log 2>&1
done
cat gpurun_out/r2c2_ab_*.log
# ncu --set full, one launch each, 4K L=3 K=5 after 300 iterations: tiled kernel and row-walking kernel (96 registers)
QGMAP_ITER=tile timeout 900 ncu --set full --clock-control none --import-source on -k regex:qgmap_iter -s 300 -c 1 -o gpurun_out/r2c2_4k_tile -f python scripts/profile_target.py full 3 5 2160 3840 2 300 > gpurun_out/r2c2_ncu_4k_tile.log 2>&1
QGMAP_LIB_PATH=build/libqgmap_w20.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:qgmap_walk -s 300 -c 1 -o gpurun_out/r2c2_4k_w20 -f python scripts/profile_target.py full 3 5 2160 3840 2 300 > gpurun_out/r2c2_ncu_4k_w20.log 2>&1
tail -2 gpurun_out/r2c2_ncu_4k_*.log; ls -la gpurun_out/*.ncu-rep | tail -3
