#!/bin/bash
# round 2, 4-GPU call (v13 kernel): bench at N=4 with the driver's defaults
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29714 bench.py --gpus 4 --e2e-its 9000 > gpurun_out/r2h4_bench_n4.json 2> gpurun_out/r2h4_bench_n4.err; echo "bench N=4 exit $?"
tail -3 gpurun_out/r2h4_bench_n4.err; cut -c1-300 gpurun_out/r2h4_bench_n4.json
