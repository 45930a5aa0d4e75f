#!/bin/bash
# round 2, GPU call 4: v10 (5x5 tap window for narrow beliefs) parity + A/B; EPE gate probe; 4K trajectory
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2c4_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c4_pytest.log
tail -5 gpurun_out/r2c4_pytest.log
CASES="full:480:640:3:5:2000,full:480:640:3:5:6000,full:480:640:3:5:0,full:2160:3840:3:5:300,full:2160:3840:3:5:3000,full:388:584:1:3:1000,full:480:640:2:9:4000,full:480:640:2:9:0"
timeout 900 python scripts/ab2.py v10 "$CASES" "tile=" > gpurun_out/r2c4_ab_v10.log 2>&1
cat gpurun_out/r2c4_ab_v10.log
timeout 900 python scripts/epe_gate_probe.py 20000 60 RubberWhale,Venus,Grove2 2 5 > gpurun_out/r2c4_epe_gate.log 2>&1
cat gpurun_out/r2c4_epe_gate.log
timeout 600 python scripts/traj_bench.py full 3 5 2160 3840 8000 1000 > gpurun_out/r2c4_traj4k.log 2>&1
cat gpurun_out/r2c4_traj4k.log
