#!/bin/bash
# last GPU call of round 2: the full-size Urban2 super-pixel case against the executed source, smoke(), and 5 resident CTAs at 48 registers (K=5)
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_refsrc_parity.py -q -m gpu -k "reference_data and urban2" -s 2>&1 | tail -6 | tee gpurun_out/r2_last.txt
timeout 60 python __graft_entry__.py smoke 2>&1 | grep smoke | tee -a gpurun_out/r2_last.txt
QGMAP_LIB_PATH=build/libqgmap_k5m5.so timeout 30 python scripts/ab3.py k5m5 2160 3840 3 5 300 2>&1 | tail -2 | tee -a gpurun_out/r2_last.txt
