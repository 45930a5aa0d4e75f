#!/bin/bash
# round 2, GPU call 1: parity suite on the row-walking kernel, then A/B against the tiled kernel and register/strip variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2c1_smi.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c1_pytest.log
timeout 600 python scripts/ab2.py cur "" "walk=;tile=QGMAP_ITER:tile;r1=QGMAP_STRIP_ROWS:1;r2=QGMAP_STRIP_ROWS:2;r4=QGMAP_STRIP_ROWS:4;r8=QGMAP_STRIP_ROWS:8;r16=QGMAP_STRIP_ROWS:16" > gpurun_out/r2c1_ab_cur.log 2>&1
for t in w28 w24 w20; do
  QGMAP_LIB_PATH=build/libqgmap_$t.so timeout 600 python scripts/ab2.py $t "" "walk=;r4=QGMAP_STRIP_ROWS:4;r8=QGMAP_STRIP_ROWS:8" > gpurun_out/r2c1_ab_$t.log 2>&1
done
tail -3 gpurun_out/r2c1_pytest.log; cat gpurun_out/r2c1_ab_*.log
