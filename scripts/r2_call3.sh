#!/bin/bash
# round 2, GPU call 3: tiled kernel v9 (packed edge epilogues, clamp-free inside path, fused band exchange) -- parity suite + A/B
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c3_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c3_pytest.log
tail -5 gpurun_out/r2c3_pytest.log
CASES="full:480:640:3:5:2000,full:480:640:3:5:0,full:2160:3840:3:5:300,full:2160:3840:3:5:3000,full:388:584:1:3:1000,full:480:640:2:9:4000,full:480:640:2:9:0"
timeout 900 python scripts/ab2.py v9 "$CASES" "tile=" > gpurun_out/r2c3_ab_v9.log 2>&1
QGMAP_LIB_PATH=build/libqgmap_t3.so timeout 900 python scripts/ab2.py t3 "$CASES" "tile=" > gpurun_out/r2c3_ab_t3.log 2>&1
QGMAP_LIB_PATH=build/libqgmap_w16.so timeout 900 python scripts/ab2.py w16 "full:480:640:3:5:2000,full:2160:3840:3:5:3000" "walk=QGMAP_ITER:walk" > gpurun_out/r2c3_ab_w16.log 2>&1
cat gpurun_out/r2c3_ab_*.log
