#!/bin/bash
# Per-SM tile queues, counters padded to one per 128-byte line and no pre-read of the SM's own queue (second attempt)
mkdir -p gpurun_out
: > gpurun_out/r2_smq2_ab.txt
QGMAP_SMQ=1 python scripts/ab3.py smq-pad 2160 3840 3 5 300 >> gpurun_out/r2_smq2_ab.txt 2>&1
QGMAP_SMQ=1 python scripts/ab3.py smq-pad 480 640 3 5 6000 >> gpurun_out/r2_smq2_ab.txt 2>&1
cat gpurun_out/r2_smq2_ab.txt
