#!/bin/bash
# Per-SM tile queues (QGMAP_SMQ=1, qg_smq_claim): co-resident CTAs take neighbouring tiles and the L components of a tile instead of tiles ~148 apart
mkdir -p gpurun_out
: > gpurun_out/r2_smq_ab.txt
python scripts/ab3.py t8 2160 3840 3 5 300 >> gpurun_out/r2_smq_ab.txt 2>&1
QGMAP_SMQ=1 python scripts/ab3.py smq 2160 3840 3 5 300 >> gpurun_out/r2_smq_ab.txt 2>&1
python scripts/ab3.py t8 480 640 3 5 6000 >> gpurun_out/r2_smq_ab.txt 2>&1
QGMAP_SMQ=1 python scripts/ab3.py smq 480 640 3 5 6000 >> gpurun_out/r2_smq_ab.txt 2>&1
QGMAP_SMQ=1 python scripts/ab3.py smq 480 640 3 5 0 >> gpurun_out/r2_smq_ab.txt 2>&1
cat gpurun_out/r2_smq_ab.txt
