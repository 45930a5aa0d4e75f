#!/bin/bash
# round 2, 8-GPU call (v13 kernel): bench at N=8 with the driver's defaults
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2h8_smi.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 > gpurun_out/r2h8_bench_n8.json 2> gpurun_out/r2h8_bench_n8.err; echo "bench N=8 exit $?"
tail -3 gpurun_out/r2h8_bench_n8.err; cut -c1-300 gpurun_out/r2h8_bench_n8.json
