#!/bin/bash
# round 2, GPU call 6: v12 (wide path by ballot, narrow path removed) parity incl. EPE gates; wide-reach sweep
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2c6_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c6_pytest.log
tail -5 gpurun_out/r2c6_pytest.log
CASES="full:480:640:3:5:0:g,full:480:640:3:5:1000:g,full:480:640:3:5:2000:g,full:480:640:3:5:6000:g,full:2160:3840:3:5:300:g,full:480:640:2:9:0:g,full:480:640:2:9:6000:g"
timeout 1500 python scripts/ab2.py v12 "$CASES" "r1=;r0.5=QGMAP_WIDE_REACH:0.5;r1.5=QGMAP_WIDE_REACH:1.5;r2.5=QGMAP_WIDE_REACH:2.5;f32=QGMAP_TAPS:f32" > gpurun_out/r2c6_ab_v12.log 2>&1
cat gpurun_out/r2c6_ab_v12.log
