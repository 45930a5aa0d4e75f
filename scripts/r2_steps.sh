#!/bin/bash
# tile steps per CTA (QGMAP_TILE_STEPS): parity at S=3 on the small problems, then A/B timing
mkdir -p gpurun_out
QGMAP_TILE_STEPS=3 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_bands.py tests/test_gpu_walk.py tests/test_golden.py tests/test_refsrc_parity.py -x -q -m gpu > gpurun_out/r2_steps_pytest.log 2>&1; tail -5 gpurun_out/r2_steps_pytest.log
QGMAP_TILE_STEPS=2 timeout 900 python -m pytest tests/test_gpu_full_size.py -x -q -m gpu > gpurun_out/r2_steps_pytest_full.log 2>&1; tail -5 gpurun_out/r2_steps_pytest_full.log
python scripts/ab2.py steps "full:2160:3840:3:5:300:g" "s1=QGMAP_TILE_STEPS:1;s2=QGMAP_TILE_STEPS:2;s3=QGMAP_TILE_STEPS:3;s4=QGMAP_TILE_STEPS:4;s6=QGMAP_TILE_STEPS:6;s10=QGMAP_TILE_STEPS:10" > gpurun_out/r2_steps_ab.txt 2>&1
python scripts/ab2.py steps "full:480:640:3:5:6000:g,full:480:640:3:5:0:g,full:388:584:1:3:1000:g,full:480:640:2:9:4000:g,full:1080:1920:3:5:300:g" "s1=QGMAP_TILE_STEPS:1;s2=QGMAP_TILE_STEPS:2;s3=QGMAP_TILE_STEPS:3;s4=QGMAP_TILE_STEPS:4" >> gpurun_out/r2_steps_ab.txt 2>&1
cat gpurun_out/r2_steps_ab.txt
