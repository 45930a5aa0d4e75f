#!/bin/bash
# programmatic dependent launch of the iteration kernel (QGMAP_PDL=1): A/B timing + parity under PDL; new per-element parity gates
mkdir -p gpurun_out
python scripts/ab2.py pdl "full:388:584:1:3:1000:g,super:480:640:3:5:3000:g,full:480:640:3:5:6000:g,full:480:640:3:5:0:g,full:480:640:2:9:4000:g,full:2160:3840:3:5:300:g" "base=;pdl=QGMAP_PDL:1" > gpurun_out/r2_pdl_ab.txt 2>&1
cat gpurun_out/r2_pdl_ab.txt
QGMAP_PDL=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_bands.py tests/test_gpu_walk.py -x -q -m gpu > gpurun_out/r2_pdl_pytest.log 2>&1; tail -3 gpurun_out/r2_pdl_pytest.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_refsrc_parity.py tests/test_golden.py -q -m gpu -s > gpurun_out/r2_parity_new.log 2>&1; grep -c "per-element" gpurun_out/r2_parity_new.log; grep "per-element" gpurun_out/r2_parity_new.log | sort -k9 -n | tail -12; tail -15 gpurun_out/r2_parity_new.log
