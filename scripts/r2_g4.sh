#!/bin/bash
# four-lane kernel without a halo warp (single edge call site, second trip after the barrier): GPU suite subset (both lane mappings are
# parametrised in the tests) + A/B against the halo-warp build; one-thread kernel without a halo warp (nh7 / nh8 builds): timing only
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bands.py tests/test_gpu_edge_cases.py tests/test_golden.py tests/test_refsrc_parity.py tests/test_gpu_full_size.py tests/test_gpu_epe.py -q -m gpu > gpurun_out/r2_g4_pytest.log 2>&1; tail -5 gpurun_out/r2_g4_pytest.log
C="super:480:640:3:5:3000:g,super:480:640:3:5:0:g,super:1080:1920:3:5:600:g,super:388:584:3:5:2000:g"
: > gpurun_out/r2_g4_ab.txt
python scripts/ab2.py nohalo "$C" "nohalo=" >> gpurun_out/r2_g4_ab.txt 2>&1
QGMAP_LIB_PATH=build/libqgmap_g4halo.so python scripts/ab2.py halo "$C" "halo=" >> gpurun_out/r2_g4_ab.txt 2>&1
C="full:2160:3840:3:5:300:g,full:480:640:3:5:0:g,full:480:640:3:5:6000:g,full:388:584:1:3:1000:g,full:480:640:2:9:4000:g,full:480:640:2:9:0:g"
python scripts/ab2.py main "$C" "main=" >> gpurun_out/r2_g4_ab.txt 2>&1
for v in nh7 nh8; do QGMAP_LIB_PATH=build/libqgmap_$v.so python scripts/ab2.py $v "$C" "$v=" >> gpurun_out/r2_g4_ab.txt 2>&1; done
sort -k3,8 -s gpurun_out/r2_g4_ab.txt
