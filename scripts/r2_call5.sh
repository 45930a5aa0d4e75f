#!/bin/bash
# round 2, GPU call 5: v11 (fp16 one-sector layout for wide beliefs on grey-level frames) parity + A/B; EPE gate probe; bench smoke
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2c5_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c5_pytest.log
tail -5 gpurun_out/r2c5_pytest.log
CASES="full:480:640:3:5:0:g,full:480:640:3:5:2000:g,full:480:640:3:5:6000:g,full:2160:3840:3:5:300:g,full:480:640:2:9:0:g,full:480:640:2:9:6000:g,full:388:584:1:3:1000:g"
timeout 1200 python scripts/ab2.py v11 "$CASES" "all=;f32taps=QGMAP_TAPS:f32;nonarrow=QGMAP_NARROW:0;neither=QGMAP_TAPS:f32,QGMAP_NARROW:0" > gpurun_out/r2c5_ab_v11.log 2>&1
cat gpurun_out/r2c5_ab_v11.log
timeout 900 python scripts/epe_gate_probe.py 20000 60 RubberWhale,Venus,Grove2 2 5 > gpurun_out/r2c5_epe_gate.log 2>&1
cat gpurun_out/r2c5_epe_gate.log
timeout 900 python bench.py --e2e-its 3000 --burnin 1000 --batch-burnin 1000 --cpu-budget 8 > gpurun_out/r2c5_bench.json 2> gpurun_out/r2c5_bench.err; echo "bench exit $?"
tail -3 gpurun_out/r2c5_bench.err; cat gpurun_out/r2c5_bench.json
