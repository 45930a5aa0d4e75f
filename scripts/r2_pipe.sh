#!/bin/bash
# software-pipelined wide-belief quadrature (QG_WIDE_PIPE): bit-identity + parity, then A/B against the same build without it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_walk.py tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_bands.py -x -q -m gpu > gpurun_out/r2_pipe_pytest.log 2>&1; tail -4 gpurun_out/r2_pipe_pytest.log
C="full:2160:3840:3:5:300:g,full:480:640:3:5:0:g,full:480:640:3:5:1000:g,full:480:640:3:5:6000:g,full:480:640:2:9:0:g,full:388:584:1:3:0:g"
python scripts/ab2.py pipe "$C" "pipe=" > gpurun_out/r2_pipe_ab.txt 2>&1
QGMAP_LIB_PATH=build/libqgmap_nopipe.so python scripts/ab2.py nopipe "$C" "nopipe=" >> gpurun_out/r2_pipe_ab.txt 2>&1
cat gpurun_out/r2_pipe_ab.txt
